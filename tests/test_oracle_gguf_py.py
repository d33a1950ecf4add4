"""The oracle's block de-quantisers against the reference's own known-answer machinery for the hot path's data formats:
`gguf.quants.dequantize` of the reference's gguf-py, which its tests/test_quants.py pins bit-exact against libggml
(SURVEY section 4).  The Python reference's outputs on seeded blocks are committed in tests/golden/gguf_py.npz
(generator: tests/golden/make_golden_gguf_py.py); get_embedding is the operator of the path that de-quantises rows
(ggml_wrapper.cpp:181-211 -> dequantize_row_q4_0 / q8_0 / q4_K / q6_K)."""
import os

import numpy as np
import pytest

from powerserve_b200 import synth
from tests import _libs as L
from tests.golden.cases import GGUF_PY_CASES, Q5K_DOT_CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gguf_py.npz")


@pytest.mark.parametrize("case", GGUF_PY_CASES, ids=lambda c: f"{c[0]}-{c[5]}")
def test_oracle_dequant_matches_gguf_py(case):
    name, t, blk, rows, n_blocks, seed, scale = case
    gold = np.load(GOLD)[f"{name}/{seed}"]
    dim = blk * n_blocks
    w = synth.random_blocks(np.random.default_rng(seed), t, rows, dim, scale)
    toks = np.arange(rows, dtype=np.int32)
    out = np.zeros(rows * dim, np.float32)
    L.oracle().ps_or_get_embedding(L.fptr(out), L.vptr(np.ascontiguousarray(w).reshape(-1)), t, dim, L.iptr(toks), rows)
    L.assert_bit_equal(out.reshape(rows, dim), gold.view(np.float32), f"oracle dequantize_row {name} vs gguf-py")


@pytest.mark.parametrize("case", Q5K_DOT_CASES, ids=lambda c: f"K{c[0]}-{c[2]}")
def test_oracle_q5k_matmul_matches_compiled_reference_golden(case):
    """ps_or_matmul on Q5_K rows (quantize_row_q8_K + the AVX2 restatement of ggml_vec_dot_q5_K_q8_K) against the committed
    outputs of the compiled reference's own kernel (tests/golden/make_golden_q5k.py)."""
    from tests.golden.make_golden_q5k import inputs
    K, rows, seed = case
    gold = np.load(os.path.join(os.path.dirname(GOLD), "q5k_dot.npz"))[f"{K}/{rows}/{seed}"]
    w, x = inputs(K, rows, seed)
    out = np.zeros(rows, np.float32)
    L.oracle().ps_or_matmul(L.Q5_K, L.vptr(np.ascontiguousarray(w).reshape(-1)), K, rows, L.fptr(x), 1, L.fptr(out))
    L.assert_bit_equal(out, gold.view(np.float32), "oracle Q5_K mat-vec vs the compiled reference")
