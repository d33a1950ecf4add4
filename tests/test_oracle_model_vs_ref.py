"""Model-level pin: the oracle's whole forward pass (prefill chunks + greedy decode) must reproduce the compiled
reference's LOGITS BIT FOR BIT and therefore its greedy token ids, for every model family / quant type of the
hot path (llama NORM-rope Q4_K, qwen2 NEOX-rope+bias Q4_0, Q8_0, mixed Q4_K+Q6_K output, head size 128)."""
import numpy as np
import pytest

from powerserve_b200 import synth
from tests import _libs as L
from tests import _model as M

pytestmark = pytest.mark.skipif(not L.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("preset", ["tiny-llama", "tiny-qwen2", "tiny-q8", "tiny-mixed", "tiny-llama-hs128"])
@pytest.mark.parametrize("n_prompt,batch", [(1, 128), (20, 8), (41, 128)])
def test_logits_and_ids_bit_exact(preset, n_prompt, batch):
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    prompt = synth.random_prompt(shape.vocab_size, n_prompt, seed=7 + n_prompt)
    n_dec = 12
    ids_ref, lg_ref, _ = M.run_reference(d, prompt, n_dec, batch_size=batch, n_threads=3, dump_logits=n_dec)
    om = M.OracleModel(d)
    ids, lg = om.generate(prompt, n_dec, batch_size=batch)
    om.close()
    L.assert_bit_equal(lg, lg_ref, f"{preset} logits")
    assert ids == ids_ref


def test_reference_is_thread_count_invariant():
    """every output element is one thread's vec_dot, so 1 vs 5 threads give identical logits."""
    d = M.model_dir("tiny-llama")
    prompt = synth.random_prompt(1024, 17)
    a = M.run_reference(d, prompt, 4, n_threads=1, dump_logits=4)
    b = M.run_reference(d, prompt, 4, n_threads=5, dump_logits=4)
    L.assert_bit_equal(a[1], b[1], "threads")
    assert a[0] == b[0]
