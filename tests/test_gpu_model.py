"""GPU parity, model level: whole forward passes through the C ABI (`ps_cuda_forward`, HOST token / logits buffers)
must reproduce the oracle's logits BIT FOR BIT — and therefore the reference's greedy token ids — for every model
family of the hot path; plus, when oracle/_ref travelled to this box, directly against the compiled reference."""
import numpy as np
import pytest

from powerserve_b200 import capi, synth
from tests import _libs as L
from tests import _model as M

pytestmark = pytest.mark.gpu

PRESETS = ["tiny-llama", "tiny-qwen2", "tiny-q8", "tiny-mixed", "tiny-llama-hs128"]


@pytest.mark.parametrize("preset", PRESETS)
@pytest.mark.parametrize("n_prompt,batch", [(1, 128), (20, 8), (41, 128)])
def test_logits_bit_exact_vs_oracle(preset, n_prompt, batch):
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    prompt = synth.random_prompt(shape.vocab_size, n_prompt, seed=7 + n_prompt)
    om = M.OracleModel(d)
    ids_o, lg_o = om.generate(prompt, 12, batch_size=batch)
    om.close()
    cm = capi.CudaModel(d, max_batch=128)
    ids_c, lg_c = cm.generate(prompt, 12, batch_size=batch)
    L.assert_bit_equal(lg_c, lg_o, f"{preset} logits")
    assert ids_c == ids_o
    # device-side greedy loop (no host round trip per token) must give the same ids
    cm.reset()
    cm.prefill(prompt, batch)
    ids_d = cm.decode_greedy(int(prompt[-1]), 12)
    assert list(ids_d) == ids_o
    cm.close()


@pytest.mark.skipif(not L.have_ref(), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("preset", ["tiny-llama", "tiny-qwen2"])
def test_ids_equal_compiled_reference(preset):
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    prompt = synth.random_prompt(shape.vocab_size, 33, seed=3)
    ids_ref, lg_ref, _ = M.run_reference(d, prompt, 32, batch_size=16, n_threads=2, dump_logits=32)
    cm = capi.CudaModel(d, max_batch=16)
    ids_c, lg_c = cm.generate(prompt, 32, batch_size=16)
    cm.close()
    L.assert_bit_equal(lg_c, lg_ref, "logits vs compiled reference")
    assert ids_c == ids_ref


def test_kv_rollback_and_batch_equivalence():
    """Prefill in one 24-token batch == 24 single-token forwards (bit-exact KV), and rollback re-decodes identically."""
    d = M.model_dir("tiny-llama")
    prompt = synth.random_prompt(1024, 25, seed=9)
    cm = capi.CudaModel(d, max_batch=32)
    cm.reset(); cm.forward(prompt[:24], lm_head=False)
    a = cm.forward([int(prompt[24])])[0]
    cm.reset()
    for t in prompt[:24]:
        cm.forward([int(t)], lm_head=False)
    b = cm.forward([int(prompt[24])])[0]
    L.assert_bit_equal(a, b, "batched vs sequential prefill")
    cm.be.kv_rollback(1)
    c = cm.forward([int(prompt[24])])[0]
    L.assert_bit_equal(a, c, "rollback")
    cm.close()


@pytest.mark.parametrize("preset", ["tiny-llama", "tiny-llama-hs128", "tiny-qwen2-r7", "tiny-q8-r3"])
@pytest.mark.parametrize("n_prompt,batch", [(76, 37), (100, 17), (131, 64)])
def test_tiled_batch_attention_ragged_shapes(preset, n_prompt, batch):
    """The register-tiled scores / P.V kernels (reduce-scatter reductions, permuted cache rows, 8 x 8 accumulator tiles)
    on shapes that hit every edge: query blocks that are not full (batch % 8), cache lengths with leftovers (n_kv % 32),
    fewer than 32 cached positions, head sizes 64 / 128, 3 / 4 / 7 query heads per kv head.  Bit-equal to the oracle and
    to the round-1 kernels (option attn_tile = 0)."""
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    prompt = synth.random_prompt(shape.vocab_size, n_prompt, seed=11 + n_prompt)
    om = M.OracleModel(d)
    ids_o, lg_o = om.generate(prompt, 4, batch_size=batch)
    om.close()
    cm = capi.CudaModel(d, max_batch=128)
    ids_c, lg_c = cm.generate(prompt, 4, batch_size=batch)
    L.assert_bit_equal(lg_c, lg_o, f"{preset} logits, tiled attention")
    assert ids_c == ids_o
    cm.reset()
    cm.be.set_option("attn_tile", 0)
    ids_b, lg_b = cm.generate(prompt, 4, batch_size=batch)
    L.assert_bit_equal(lg_b, lg_c, f"{preset} logits, round-1 batch attention kernels")
    cm.close()


@pytest.mark.parametrize("preset", ["tiny-hs32", "tiny-hs96", "tiny-hs256"])
def test_head_sizes_off_the_templated_paths(preset):
    """Head sizes 32 / 96 / 256: the decode scores kernel's generic step loop (ST = 8 template), the head-size-templated
    batch scores kernel for 32 and 256, the round-1 batch kernels for 96.  Prefill in chunks, host-driven decode and the
    device-resident greedy loop, bit-equal to the oracle."""
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    prompt = synth.random_prompt(shape.vocab_size, 45, seed=5)
    om = M.OracleModel(d)
    ids_o, lg_o = om.generate(prompt, 8, batch_size=19)
    om.close()
    cm = capi.CudaModel(d, max_batch=32)
    ids_c, lg_c = cm.generate(prompt, 8, batch_size=19)
    L.assert_bit_equal(lg_c, lg_o, f"{preset} logits")
    assert ids_c == ids_o
    cm.reset()
    cm.prefill(prompt, 19)
    assert list(cm.decode_greedy(int(prompt[-1]), 8)) == ids_o
    cm.close()
