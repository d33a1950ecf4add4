"""GPU parity at the BENCHMARKED context length and beyond (VERDICT r01 "parity holes"): 2100- and 4400-token prompts
(tcgen05 prefill GEMM in chunks of 128, batched attention) followed by fused / graph-replayed decode steps, on the real
per-layer shapes of Llama-3.1-8B and Llama-3.2-1B, against the committed golden vectors of the compiled reference
(tests/golden/long_ctx.npz) — bit for bit.  n_ctx = 4096 is the bench configuration; n_ctx = 8192 puts the decode
attention beyond the range where the probabilities / V^T rows of a kv group are shared-memory resident."""
import os

import numpy as np
import pytest

from powerserve_b200 import capi, gguf, synth
from tests import _libs as L
from tests import _model as M
from tests.golden import cases
from tests.golden.make_golden_long import long_prompt

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden", "long_ctx.npz")


def _model(preset, n_ctx, max_batch=128):
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    g = gguf.GGUFFile(os.path.join(d, "ggml", "weights.gguf"))
    desc = capi.desc_from_model_json(synth.model_json(shape), max_batch=max_batch, n_ctx=n_ctx)
    return capi.CudaModel(desc=desc, tensors=g.tensors), g


@pytest.mark.parametrize("persist", [0, 1], ids=["per-phase-kernels", "persistent-step-kernel"])
@pytest.mark.parametrize("n_ctx", [4096, 8192])
@pytest.mark.parametrize("case", cases.LONG_CASES, ids=lambda c: f"{c[0]}-{c[1]}")
def test_long_context_decode_matches_reference(case, n_ctx, persist):
    preset, n_prompt, batch, n_dec = case
    if n_prompt + n_dec + 1 > n_ctx:
        pytest.skip("prompt does not fit this n_ctx")
    gold = np.load(G)
    key = f"{preset}/{n_prompt}/{batch}"
    cm, keep = _model(preset, n_ctx)
    cm.be.set_option("persist", persist)
    prompt = long_prompt(preset, n_prompt)
    ids, lg = cm.generate(prompt, n_dec, batch_size=batch)              # host-driven steps: logits of every step
    assert ids == list(gold[key + "/ids"])
    assert (L.bits(lg) == gold[key + "/logits_bits"]).all(), "logits differ bitwise from the compiled reference"
    cm.reset(); cm.prefill(prompt, batch)
    dev_ids = list(cm.decode_greedy(int(prompt[-1]), n_dec))             # device-resident greedy loop (graph replay)
    assert dev_ids == ids
    assert cm.be.counter("tc_error") == 0 and cm.be.counter("step_error") == 0
    assert (cm.be.counter("step_launches") > 0) == bool(persist)
    if not persist:   # the opt-in group-synchronised attention kernel (applies at n_ctx 4096; falls back to the two kernels beyond)
        cm.be.set_option("attn_group", 1)
        for _ in range(3):   # repeated: its cross-CTA hand-offs are timing dependent
            ids2, lg2 = cm.generate(prompt, n_dec, batch_size=batch)
            assert ids2 == ids and (L.bits(lg2) == gold[key + "/logits_bits"]).all(), "group-synchronised attention differs from the compiled reference"
        assert cm.be.counter("step_error") == 0
    cm.close()
