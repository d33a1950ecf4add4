"""GPU tests of the C-ABI entries the round-1 suite did not touch: ps_cuda_copy_2d (the `copy` / `cont` twin), the KV
position bookkeeping (truncate / advance / rollback, kv_cache.hpp:97-163 semantics), the exposed cache pointers
(ps_cuda_kv_k / ps_cuda_kv_v against the oracle's cache contents) and ps_cuda_logits_dev."""
import numpy as np
import pytest

from powerserve_b200 import capi, synth
from tests import _libs as L
from tests import _model as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    b = capi.CudaBackend(capi.ModelDesc(512, 1536, 1, 8, 2, 64, 1024, 512, 1e-5, 64, 0, 5e5, 1.0, 1.0, 0, 32, 0, 1))
    yield b
    b.close()


@pytest.mark.parametrize("ne0,ne1", [(64, 8), (37, 5), (1, 129), (128, 1)])
def test_copy_2d_strided(be, ne0, ne1):
    rng = np.random.default_rng(ne0 * 131 + ne1)
    # (a) the V-cache store of NormAttention::build: contiguous {ne0, ne1} source -> transposed destination with row pitch n_ctx
    n_ctx = 160
    src = rng.standard_normal(ne0 * ne1).astype(np.float32)
    dst0 = rng.standard_normal(n_ctx * ne0).astype(np.float32)
    sd, dd = be.upload(src), be.upload(dst0)
    be.copy_2d(dd, 4 * n_ctx, 4, sd, 4, 4 * ne0, ne0, ne1)          # dst[i0][i1] (pitch n_ctx) = src[i1][i0]
    want = dst0.reshape(ne0, n_ctx).copy()
    want[:, :ne1] = src.reshape(ne1, ne0).T
    L.assert_bit_equal(dd.numpy(), want.reshape(-1), "transposed strided store")
    # (b) `cont` of a permuted view: {ne0, ne1} read with swapped strides into a contiguous buffer
    cd = be.empty(ne0 * ne1)
    be.copy_2d(cd, 4, 4 * ne1, sd, 4 * ne0, 4, ne1, ne0)           # dst[j][i] contiguous in i = src[i][j]
    L.assert_bit_equal(cd.numpy(), src.reshape(ne1, ne0).T.reshape(-1), "cont of a permuted view")
    for b in (sd, dd, cd):
        b.free()


def test_kv_bookkeeping_and_cache_contents():
    d = M.model_dir("tiny-llama")
    shape = synth.PRESETS["tiny-llama"]
    prompt = synth.random_prompt(shape.vocab_size, 21, seed=3)
    cm = capi.CudaModel(d, max_batch=8)
    om = M.OracleModel(d)
    cm.prefill(prompt, 8); om.reset()
    i = 0
    while i < len(prompt) - 1:                                           # same chunking on the oracle
        bs = min(8, len(prompt) - 1 - i)
        om.forward(prompt[i:i + bs], lm_head=False)
        i += bs
    n = len(prompt) - 1
    assert cm.position == n == om.position
    kvd, n_ctx = shape.kv_dim, shape.n_ctx
    for layer in range(shape.n_layers):                                  # cache contents, layouts of ggml_kv_cache.cpp:35-58
        k = cm.be.read_device(cm.be.kv_k(layer), n * kvd)
        v = cm.be.read_device(cm.be.kv_v(layer), kvd * n_ctx).reshape(kvd, n_ctx)[:, :n]
        ko = np.ctypeslib.as_array(om.lib.ps_or_model_k_cache(om.h, layer), shape=(n_ctx * kvd,))[: n * kvd]
        vo = np.ctypeslib.as_array(om.lib.ps_or_model_v_cache(om.h, layer), shape=(kvd * n_ctx,)).reshape(kvd, n_ctx)[:, :n]
        L.assert_bit_equal(k, ko, f"K cache layer {layer}")
        L.assert_bit_equal(np.ascontiguousarray(v), np.ascontiguousarray(vo), f"V cache layer {layer}")
    assert cm.be.kv_k(shape.n_layers) is None and cm.be.kv_v(-1) is None
    # truncate_tokens (kv_cache.hpp:265-271): only ever shrinks
    cm.be.kv_truncate(n + 5); assert cm.position == n
    cm.be.kv_truncate(n - 4); assert cm.position == n - 4
    cm.be.kv_advance(4); assert cm.position == n                       # advance_tokens: the 4 slots still hold their rows
    lg_c = cm.forward([int(prompt[-1])])[0]
    lg_o = om.forward([int(prompt[-1])])[0]
    L.assert_bit_equal(lg_c, lg_o, "decode after truncate + advance")
    # ps_cuda_logits_dev: the device copy of what forward() returned
    L.assert_bit_equal(cm.be.read_device(cm.be.logits_dev(), shape.vocab_size), lg_c, "logits_dev")
    with pytest.raises(capi.PsCudaError):
        cm.be.kv_advance(n_ctx)                                          # KV full
    with pytest.raises(capi.PsCudaError):
        cm.be.kv_rollback(cm.position + 1)
    cm.close(); om.close()
