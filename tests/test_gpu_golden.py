"""GPU parity against the COMMITTED golden vectors of the compiled reference (tests/golden/*.npz): the CUDA library,
called through the C ABI, must reproduce the reference's outputs bit for bit — without the oracle in the loop."""
import os

import numpy as np
import pytest

from powerserve_b200 import capi, synth
from tests import _libs as L
from tests import _model as M
from tests.golden import cases

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


class CudaGoldenBackend:
    """the cases.* backend protocol over libps_cuda.so"""

    def __init__(self, hs=64, rope_type=0, base=5e5):
        self.be = capi.CudaBackend(capi.ModelDesc(512, 1536, 1, 8, 2, hs, 1024, 1024, 1e-5, hs, rope_type, base, 1.0, 1.0, 0, 32, 0, 1))

    def close(self):
        self.be.close()

    def matmul(self, t, w, k, n, x, bs):
        wd = self.be.register_weight(w, t, k, n)
        xd, dd = self.be.upload(x), self.be.empty(n * bs)
        self.be.matmul(dd, wd, t, k, n, xd, bs)
        out = dd.numpy().reshape(bs, n)
        xd.free(); dd.free(); self.be.unregister_weight(w)
        return out

    def rmsnorm(self, x, w, dim, bs, eps):
        xd, wd, dd = self.be.upload(x), self.be.upload(w), self.be.empty(x.size)
        self.be.rmsnorm(dd, xd, wd, dim, bs, eps)
        return dd.numpy().reshape(x.shape)

    def rope(self, x, hs, nh, bs, pos, mode, base):
        xd, dd = self.be.upload(x), self.be.empty(x.size)
        self.be.rope(dd, xd, hs, nh, bs, pos)
        return dd.numpy().reshape(x.shape)

    def softmax_ext(self, x, mask, ne0, ne1, ne2, scale):
        xd, md, dd = self.be.upload(x), self.be.upload(mask), self.be.empty(x.size)
        self.be.softmax_ext(dd, xd, md, ne0, ne1, ne2, scale)
        return dd.numpy().reshape(x.shape)

    def silu_hadamard(self, g, u):
        gd, ud, dd = self.be.upload(g), self.be.upload(u), self.be.empty(g.size)
        self.be.silu_hadamard(dd, gd, ud, g.size)
        return dd.numpy()

    def get_embedding(self, w, t, dim, vocab, tokens):
        wd = self.be.register_weight(w, t, dim, vocab)
        dd = self.be.empty(dim * len(tokens))
        self.be.get_embedding(dd, wd, t, dim, tokens)
        out = dd.numpy().reshape(len(tokens), dim)
        self.be.unregister_weight(w)
        return out


def test_cuda_matmul_matches_reference_golden():
    gold = np.load(os.path.join(G, "ops.npz"))
    be = CudaGoldenBackend()
    for k, v in cases.case_matmul(be).items():
        assert (gold[f"matmul/{k}"] == v).all(), k
    be.close()


def test_cuda_small_ops_match_reference_golden():
    gold = np.load(os.path.join(G, "ops.npz"))
    norm, neox = CudaGoldenBackend(64, 0, 5e5), CudaGoldenBackend(64, 2, 1e6)

    class Mixed:  # rope mode is a property of the context (model config), everything else is mode-independent
        def __getattr__(self, n):
            return getattr(norm, n)

        def rope(self, x, hs, nh, bs, pos, mode, base):
            return (neox if mode == 2 else norm).rope(x, hs, nh, bs, pos, mode, base)

    for k, v in cases.case_small_ops(Mixed()).items():
        assert (gold[f"small/{k}"] == v).all(), k
    norm.close(); neox.close()


@pytest.mark.parametrize("preset,n_prompt,batch,n_dec", cases.MODEL_CASES)
def test_cuda_models_match_reference_golden(preset, n_prompt, batch, n_dec):
    gold = np.load(os.path.join(G, "models.npz"))
    prompt = synth.random_prompt(synth.PRESETS[preset].vocab_size, n_prompt, seed=7 + n_prompt)
    cm = capi.CudaModel(M.model_dir(preset), max_batch=128)
    ids, lg = cm.generate(prompt, n_dec, batch_size=batch)
    cm.close()
    key = f"{preset}/{n_prompt}/{batch}"
    assert ids == list(gold[key + "/ids"])
    assert (L.bits(lg) == gold[key + "/logits_bits"]).all()
