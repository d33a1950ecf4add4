"""The arithmetic claim behind the tcgen05 prefill GEMM (powerserve_b200/csrc/ps_tc.cuh): the integer
lane sums of ggml_vec_dot_q4_K_q8_K are EXACT when computed as an fp16 x fp16 -> fp32 contraction, in any accumulation
order - every operand is an integer fp16 represents exactly, every product is exact in fp32, and every partial sum stays
below 2^24.  Pinned here on CPU with numpy (no GPU needed): worst-case bounds and random data in shuffled orders."""
import numpy as np


def test_operands_are_exact_in_fp16():
    sc, q4, q8 = np.arange(64), np.arange(16), np.arange(-128, 128)
    a = np.multiply.outer(sc, q4).reshape(-1)                     # weight operand sc_j * q4 <= 945
    assert a.max() == 945 and np.array_equal(a.astype(np.float16).astype(np.int64), a)
    assert np.array_equal(q8.astype(np.float16).astype(np.int64), q8)
    hs = np.arange(-16 * 128, 16 * 127 + 1)                        # half-sub-block sums of 16 q8 values (mins operand)
    assert np.array_equal(hs.astype(np.float16).astype(np.int64), hs)   # |.| <= 2048 = 2^11: exact in fp16


def test_worst_case_sums_stay_below_2_pow_24():
    lane = 8 * 4 * 945 * 128          # S_l: 8 sub-blocks x 4 elements, |sc*q4| <= 945, |q8| <= 128
    mins = 2 * 8 * 63 * 2048          # P: 16 half-sub-block sums, m_j <= 63
    assert lane < 2 ** 24 and mins < 2 ** 24


def test_fp32_accumulation_is_exact_in_any_order():
    rng = np.random.default_rng(0)
    for _ in range(200):
        extreme = rng.random() < 0.3
        sc = np.full(8, 63) if extreme else rng.integers(0, 64, 8)
        q4 = np.full((8, 4), 15) if extreme else rng.integers(0, 16, (8, 4))
        q8 = rng.choice([-128, 127], (8, 4)) if extreme else rng.integers(-128, 128, (8, 4))
        a = (sc[:, None] * q4).astype(np.float16).reshape(-1)      # one AVX lane's K = 32 operand row
        b = q8.astype(np.float16).reshape(-1)
        exact = int((sc[:, None].astype(np.int64) * q4 * q8).sum())
        prod = a.astype(np.float32) * b.astype(np.float32)         # fp16 x fp16 products, exact in fp32
        for _ in range(4):
            acc = np.float32(0)
            for v in prod[rng.permutation(32)]:
                acc = np.float32(acc + v)
            assert int(acc) == exact and float(acc) == float(exact)
