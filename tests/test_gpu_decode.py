"""GPU parity of the FUSED decode path (persistent TMA-fed Q4_K mat-vec with fused RMSNorm/quantise prologue and
bias/residual/SiLU epilogue, two-kernel decode attention, PDL chaining, CUDA-graph replay) against (a) the table-op
path of the same library and (b) the oracle — bit for bit, logits and greedy ids."""
import numpy as np
import pytest

from powerserve_b200 import capi, synth
from tests import _libs as L
from tests import _model as M

pytestmark = pytest.mark.gpu


def run(cm, prompt, n_dec, fused, graph, pdl=1, persist=0, attn_chunk=0, attn_fused=0, ops_graph=0, attn_group=0):
    cm.be.set_option("attn_group", attn_group)  # 1 = decode attention as ONE group-synchronised kernel per layer (opt-in), 0 = scores kernel + soft-max / P.V kernel (default)
    cm.be.set_option("ops_graph", ops_graph)  # 0 with fused = 0: the plain table-op path (one launch per op, host-fed positions)
    cm.be.set_option("fused", fused)
    cm.be.set_option("graph", graph)
    cm.be.set_option("pdl", pdl)
    cm.be.set_option("persist", persist)     # 1 = the whole step as ONE persistent kernel (ps_step.cuh)
    cm.be.set_option("attn_chunk", attn_chunk)
    cm.be.set_option("attn_fused", attn_fused)  # 1 = decode attention as one cluster kernel per layer, 0 = scores kernel + soft-max / P.V kernel
    return cm.generate(prompt, n_dec, batch_size=16)


@pytest.mark.parametrize("preset", ["tiny-llama", "tiny-llama-hs128", "slice-1b", "slice-8b"])
def test_fused_equals_table_ops_and_oracle(preset):
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    prompt = synth.random_prompt(shape.vocab_size, 37, seed=11)
    n_dec = 10
    cm = capi.CudaModel(d, max_batch=16)
    ids_u, lg_u = run(cm, prompt, n_dec, fused=0, graph=0)
    ids_f, lg_f = run(cm, prompt, n_dec, fused=1, graph=0, pdl=0)
    L.assert_bit_equal(lg_f, lg_u, f"{preset}: fused vs table ops")
    ids_p, lg_p = run(cm, prompt, n_dec, fused=1, graph=0, pdl=1)
    L.assert_bit_equal(lg_p, lg_u, f"{preset}: fused+PDL vs table ops")
    ids_g, lg_g = run(cm, prompt, n_dec, fused=1, graph=1, pdl=1)
    L.assert_bit_equal(lg_g, lg_u, f"{preset}: fused+PDL+graph vs table ops")
    assert ids_u == ids_f == ids_p == ids_g
    for g_, p_ in ((0, 0), (1, 1)):     # the group-synchronised one-kernel decode attention (the runs above use the two-kernel default)
        ids_k, lg_k = run(cm, prompt, n_dec, fused=1, graph=g_, pdl=p_, attn_group=1)
        L.assert_bit_equal(lg_k, lg_u, f"{preset}: group-synchronised attention (graph {g_}, pdl {p_}) vs table ops")
        assert ids_k == ids_u
    assert cm.be.counter("step_error") == 0
    ids_2, lg_2 = run(cm, prompt, n_dec, fused=1, graph=1, pdl=1, attn_fused=1)
    L.assert_bit_equal(lg_2, lg_u, f"{preset}: one-kernel (cluster) attention vs table ops")
    assert ids_2 == ids_u
    # device-resident greedy loop (graph replay per step, token fed back on the device)
    cm.reset(); cm.prefill(prompt, 16)
    ids_d = list(cm.decode_greedy(int(prompt[-1]), n_dec))
    assert ids_d == ids_u
    assert cm.be.counter("graph_replays") >= n_dec
    # the persistent step kernel: host-driven steps (logits) and the device-resident loop
    assert cm.be.counter("step_ok") == 1
    n0 = cm.be.counter("step_launches")
    ids_s, lg_s = run(cm, prompt, n_dec, fused=1, graph=1, persist=1)
    L.assert_bit_equal(lg_s, lg_u, f"{preset}: persistent step kernel vs table ops")
    assert ids_s == ids_u and cm.be.counter("step_launches") >= n0 + n_dec
    cm.reset(); cm.prefill(prompt, 16)
    assert list(cm.decode_greedy(int(prompt[-1]), n_dec)) == ids_u
    assert cm.be.counter("step_error") == 0
    cm.close()
    om = M.OracleModel(d)
    ids_o, lg_o = om.generate(prompt, n_dec, batch_size=16)
    om.close()
    L.assert_bit_equal(lg_u, lg_o, f"{preset}: table ops vs oracle")
    assert ids_o == ids_u


def test_fused_long_context_matches_table_ops():
    """n_kv crossing the 8- / 32-element tails and several 32-position chunks."""
    d = M.model_dir("tiny-llama")
    prompt = synth.random_prompt(1024, 130, seed=4)
    cm = capi.CudaModel(d, max_batch=64)
    ids_u, lg_u = run(cm, prompt, 40, fused=0, graph=0)
    ids_g, lg_g = run(cm, prompt, 40, fused=1, graph=1)
    L.assert_bit_equal(lg_g, lg_u, "long context")
    assert ids_g == ids_u
    ids_2, lg_2 = run(cm, prompt, 40, fused=1, graph=1, attn_fused=1)
    L.assert_bit_equal(lg_2, lg_u, "long context, one-kernel (cluster) attention")
    ids_s, lg_s = run(cm, prompt, 40, fused=1, graph=1, persist=1)
    L.assert_bit_equal(lg_s, lg_u, "long context, persistent step kernel")
    assert ids_s == ids_u
    cm.close()


def test_step_kernel_chunked_attention():
    """soft-max rows longer than the shared-memory-resident chunk (forced down to 256 positions): the three-pass path of
    the persistent kernel's attention phase, n_kv crossing chunk, 32- and 8-element boundaries."""
    d = M.model_dir("tiny-llama")
    prompt = synth.random_prompt(1024, 250, seed=9)
    cm = capi.CudaModel(d, max_batch=64)
    ids_u, lg_u = run(cm, prompt, 24, fused=0, graph=0)
    ids_s, lg_s = run(cm, prompt, 24, fused=1, graph=1, persist=1, attn_chunk=256)
    L.assert_bit_equal(lg_s, lg_u, "chunked attention")
    assert ids_s == ids_u and cm.be.counter("step_error") == 0
    cm.close()


@pytest.mark.parametrize("preset", ["slice-8b", "slice-1b", "tiny-llama"])
def test_matvec_ksplit_and_deferred_stream_are_bit_exact(preset):
    """K-split walks (several warps share a row octet; the FMA chains travel from warp to warp as a token, stage by stage in
    row order) and the deferred weight stream are scheduling choices: every setting must leave the table-op path's bits."""
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    prompt = synth.random_prompt(shape.vocab_size, 41, seed=5)
    cm = capi.CudaModel(d, max_batch=16)
    ids_u, lg_u = run(cm, prompt, 8, fused=0, graph=0)
    for ksplit, defer, kb in [(0, 0, 0), (2, 0, 0), (4, 62, 0), (8, 0, 0), (8, 62, 1), (4, 0, 2)]:
        cm.be.set_option("rw_ksplit", ksplit)
        cm.be.set_option("rw_defer", defer)
        cm.be.set_option("rw_kb", kb)        # cap on the blocks per stage = per token hop
        ids_f, lg_f = run(cm, prompt, 8, fused=1, graph=1)
        L.assert_bit_equal(lg_f, lg_u, f"{preset}: rw_ksplit={ksplit} rw_defer={defer} rw_kb={kb}")
        assert ids_f == ids_u
    cm.close()


@pytest.mark.parametrize("preset", ["tiny-qwen2", "tiny-qwen2-r7", "tiny-q8", "tiny-q8-r3", "tiny-mixed", "tiny-llama", "tiny-q5k"])
def test_graph_replayed_operator_table_step_for_every_weight_type(preset):
    """Models off the all-Q4_K fused path (Q4_0 / Q8_0 matrices, a Q6_K output matrix, NEOX rope + biases, 3 or 7 query heads per
    kv head) decode through decode_step_ops: table-op weight products + the position-from-device decode attention kernels,
    captured once and replayed per token, greedy pick on the device.  Bit-exact vs the plain table-op path and the oracle."""
    d = M.model_dir(preset)
    shape = synth.PRESETS[preset]
    prompt = synth.random_prompt(shape.vocab_size, 29, seed=13)
    n_dec = 12
    cm = capi.CudaModel(d, max_batch=16)
    ids_u, lg_u = run(cm, prompt, n_dec, fused=0, graph=0, ops_graph=0)
    ids_n, lg_n = run(cm, prompt, n_dec, fused=0, graph=0, ops_graph=1)        # the step's kernels launched one by one
    L.assert_bit_equal(lg_n, lg_u, f"{preset}: operator-table step vs table ops")
    n0 = cm.be.counter("graph_replays")
    ids_g, lg_g = run(cm, prompt, n_dec, fused=0, graph=1, ops_graph=1)        # replayed as a graph, host reads the logits
    L.assert_bit_equal(lg_g, lg_u, f"{preset}: graph-replayed operator-table step vs table ops")
    assert ids_u == ids_n == ids_g and cm.be.counter("graph_replays") >= n0 + n_dec
    cm.reset(); cm.prefill(prompt, 16)
    n0 = cm.be.counter("graph_replays")
    ids_d = list(cm.decode_greedy(int(prompt[-1]), n_dec))                     # device-resident loop: pick + feedback on the device
    assert ids_d == ids_u and cm.be.counter("graph_replays") >= n0 + n_dec
    L.assert_bit_equal(cm.be.read_device(cm.be.logits_dev(), shape.vocab_size), lg_u[-1], f"{preset}: last logits of the device loop")
    # all-Q4_0 / all-Q8_0 models: the FUSED 32-block path (ps_mv32.cuh: TMA-fed octet mat-vec with the RMSNorm / SiLU.up +
    # Q8_0 quantiser prologue and bias / residual / pick epilogue, 7 launches per layer)
    uniform32 = shape.wtype in (synth.GGML_Q4_0, synth.GGML_Q8_0) and shape.output_type in (None, shape.wtype)
    assert cm.be.counter("mv32_ok") == (1 if uniform32 else 0)
    if uniform32:
        ids_m, lg_m = run(cm, prompt, n_dec, fused=1, graph=0, pdl=0)
        L.assert_bit_equal(lg_m, lg_u, f"{preset}: fused 32-block step vs table ops")
        ids_p, lg_p = run(cm, prompt, n_dec, fused=1, graph=1, pdl=1)
        L.assert_bit_equal(lg_p, lg_u, f"{preset}: fused 32-block step, PDL + graph, vs table ops")
        cm.be.set_option("mv_kpar", 0)      # without the block-parallel mode of the launches with one octet per CTA and long rows
        ids_k, lg_k = run(cm, prompt, n_dec, fused=1, graph=1, pdl=1)
        cm.be.set_option("mv_kpar", 1)
        L.assert_bit_equal(lg_k, lg_u, f"{preset}: fused 32-block step without block-parallel launches vs table ops")
        assert ids_k == ids_u
        cm.reset(); cm.prefill(prompt, 16)
        ids_l = list(cm.decode_greedy(int(prompt[-1]), n_dec))
        assert ids_m == ids_p == ids_l == ids_u
        L.assert_bit_equal(cm.be.read_device(cm.be.logits_dev(), shape.vocab_size), lg_u[-1], f"{preset}: last logits of the fused device loop")
    cm.close()
    om = M.OracleModel(d)
    ids_o, lg_o = om.generate(prompt, n_dec, batch_size=16)
    om.close()
    L.assert_bit_equal(lg_u, lg_o, f"{preset}: table ops vs oracle")
    assert ids_u == ids_o
