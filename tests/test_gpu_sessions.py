"""Server-side batching: several independent sessions (KV sets) over one set of weights, advanced one token
each per forward pass - one weight stream for the whole batch, attention per column over its own cache.  Parity: every
session's logits are bit-identical to decoding that sequence alone with batch 1 (and hence to the oracle, which the
single-sequence tests pin)."""
import numpy as np
import pytest

from powerserve_b200 import capi, synth
from tests import _libs as L
from tests import _model as M

pytestmark = pytest.mark.gpu


def alone(cm, prompt, n_dec, batch):
    cm.be.set_option("fused", 0)
    cm.be.set_option("ops_graph", 0)
    out = cm.generate(prompt, n_dec, batch_size=batch)
    cm.be.set_option("fused", 1)
    cm.be.set_option("ops_graph", 1)
    return out


@pytest.mark.parametrize("preset,n_sess", [("tiny-llama", 3), ("tiny-qwen2", 5), ("tiny-q8-r3", 2), ("slice-1b", 16), ("tiny-mixed", 1)])
def test_batched_sessions_equal_single_sequence_decoding(preset, n_sess):
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    n_dec = 7
    cm = capi.CudaModel(d, max_batch=16)
    prompts = [synth.random_prompt(shape.vocab_size, 3 + 7 * s % 41 + s, seed=100 + s) for s in range(n_sess)]
    want = [alone(cm, p, n_dec + 3, 16) for p in prompts]             # (ids, logits [n_dec + 3][vocab]) per sequence, decoded alone
    cm.reset()
    sids = [0] + [cm.session_create() for _ in range(n_sess - 1)]     # session 0 = the context's own cache
    for sid, p in zip(sids, prompts):                                 # per-session prefill through the ordinary calls
        cm.session_select(sid)
        cm.reset()
        cm.prefill(p, 16)
        assert cm.position == len(p) - 1 == cm.session_position(sid)
    cm.session_select(sids[-1])                                       # the batch may contain the selected session or not
    toks = [int(p[-1]) for p in prompts]
    for step in range(n_dec):
        order = list(range(n_sess)) if step % 2 == 0 else list(reversed(range(n_sess)))   # column order must not matter
        lg, ids = cm.forward_sessions([sids[s] for s in order], [toks[s] for s in order])
        for col, s in enumerate(order):
            L.assert_bit_equal(lg[col], want[s][1][step], f"{preset}: session {s} step {step}")
            assert int(ids[col]) == int(np.argmax(lg[col])) == want[s][0][step]
            toks[s] = int(ids[col])
    for s, sid in enumerate(sids):
        assert cm.session_position(sid) == len(prompts[s]) - 1 + n_dec
    # a session keeps working through the single-sequence calls after batched steps (fused decode on its cache)
    cm.session_select(sids[0])
    ids_more = [int(t) for t in cm.decode_greedy(toks[0], 3)]
    assert ids_more == want[0][0][n_dec:n_dec + 3]
    cm.session_select(0)
    for sid in sids[1:]:
        cm.session_destroy(sid)
    cm.close()
