"""Golden vectors for the Q5_K dot product from the COMPILED reference (oracle/_ref/libggml_ref.so: ggml_vec_dot_q5_K_q8_K with
quantize_row_q8_K activations), so that the oracle's restatement can be checked on a box without /root/reference.
Run in the build container after `make -C oracle ref`:  python tests/golden/make_golden_q5k.py  ->  tests/golden/q5k_dot.npz
Inputs are regenerated from the seeds by the test."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from powerserve_b200 import synth  # noqa: E402
from tests import _libs as L  # noqa: E402
from tests.golden.cases import Q5K_DOT_CASES  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def inputs(K, rows, seed):
    rng = np.random.default_rng(seed)
    w = synth.random_blocks(rng, L.Q5_K, rows, K, K ** -0.5)
    x = rng.standard_normal(K).astype(np.float32)
    return w, x


def main():
    r, o = L.ref_ggml(), L.oracle()
    out = {}
    for K, rows, seed in Q5K_DOT_CASES:
        w, x = inputs(K, rows, seed)
        xq = np.zeros(o.ps_or_row_size(L.Q8_K, K), np.uint8)
        r.quantize_row_q8_K(L.fptr(x), L.vptr(xq), K)
        res = np.zeros(rows, np.float32)
        for n in range(rows):
            one = np.zeros(1, np.float32)
            row = np.ascontiguousarray(w[n])
            r.ggml_vec_dot_q5_K_q8_K(K, L.fptr(one), 0, L.vptr(row), 0, L.vptr(xq), 0, 1)
            res[n] = one[0]
        out[f"{K}/{rows}/{seed}"] = res.view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "q5k_dot.npz"), **out)
    print("wrote", os.path.join(HERE, "q5k_dot.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
