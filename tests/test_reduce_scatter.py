"""Host-side model of the reduce-scatter reductions used by the attention kernels (csrc/ps_kernels.cuh:
ps_k_attn_scores_tile, ps_k_attn_pv_tile / ps_rs_stage; csrc/ps_decode.cuh: ps_k_attn1).

The reference reduces the 32 per-lane fp32 partial sums of a dot product with GGML_F32x8_REDUCE, which on a warp is the
butterfly xor 16, 8, 4, 1, 2 (libs/ggml/src/ggml.c:2092-2131 and the F32x8 reduce macro).  The kernels hold several
independent sums per lane and let every lane keep only half of them per stage.  These tests replay both schemes lane by lane
in numpy float32 (IEEE add, same rounding as __fadd_rn) and check that
  * the value a lane ends up with is BIT-identical to what the full butterfly leaves for that sum, and
  * it sits in the lane / slot the kernels store from (the index maps in the kernels' comments).
No GPU involved: this pins the index algebra, the GPU tests pin the kernels."""
import numpy as np
import pytest

STAGES = (16, 8, 4, 1, 2)


def butterfly(parts: np.ndarray) -> np.float32:
    """parts[32] -> the value every lane holds after the full butterfly."""
    v = parts.astype(np.float32).copy()
    for st in STAGES:
        v = (v + v[np.arange(32) ^ st]).astype(np.float32)
    assert len({x.tobytes() for x in v}) == 1      # commutativity: all lanes agree bit for bit
    return v[0]


def rand_parts(rng, *shape):
    # mixed magnitudes and signs so that the order of the adds matters
    return (rng.standard_normal(shape) * np.exp(rng.uniform(-8, 8, shape))).astype(np.float32)


def shfl_xor(vals, mask):
    """vals[lane] -> what each lane receives from lane ^ mask"""
    return vals[np.arange(32) ^ mask]


def test_scores_tile_eight_rows_permuted_slots():
    """ps_k_attn_scores_tile: slot t of lane l holds cache row t ^ (l >> 2); 4 + 2 + 1 + 1 + 1 shuffles; slot 0 ends as row l >> 2."""
    rng = np.random.default_rng(1)
    part = rand_parts(rng, 32, 8)                   # part[lane][row]: the lane's FMA chain for that row
    ref = [butterfly(part[:, r]) for r in range(8)]
    c = np.arange(32) >> 2
    s = np.stack([part[np.arange(32), t ^ c] for t in range(8)], axis=1)   # s[lane][slot]
    n_shfl = 0
    for t in range(4):
        s[:, t] = s[:, t] + shfl_xor(s[:, t + 4], 16); n_shfl += 1
    for t in range(2):
        s[:, t] = s[:, t] + shfl_xor(s[:, t + 2], 8); n_shfl += 1
    s[:, 0] = s[:, 0] + shfl_xor(s[:, 1], 4); n_shfl += 1
    s[:, 0] = s[:, 0] + shfl_xor(s[:, 0], 1); n_shfl += 1
    s[:, 0] = s[:, 0] + shfl_xor(s[:, 0], 2); n_shfl += 1
    assert n_shfl == 9
    for lane in range(32):
        assert s[lane, 0].tobytes() == ref[lane >> 2].tobytes(), lane


@pytest.mark.parametrize("R2", [1, 2, 4, 8])
def test_decode_scores_rows_and_heads(R2):
    """ps_k_attn1<R2, ST>: slot (t, hh) of a lane holds row t ^ ct and head hh ^ ch; stages 16 / 8 / 4 halve the rows, stages
    1 / 2 halve the heads (when there are heads left to halve)."""
    HB = {1: 0, 2: 1, 4: 2, 8: 3}[R2]
    rng = np.random.default_rng(10 + R2)
    part = rand_parts(rng, 32, 8, R2)               # part[lane][row][head]
    ref = [[butterfly(part[:, r, h]) for h in range(R2)] for r in range(8)]
    lanes = np.arange(32)
    ct = lanes >> 2
    ch = ((lanes & 1) << (HB - 1) if HB >= 1 else 0) | (((lanes >> 1) & 1) << (HB - 2) if HB >= 2 else 0)
    ch = np.broadcast_to(ch, (32,))
    s = np.empty((32, 8, R2), np.float32)
    for t in range(8):
        for hh in range(R2):
            s[:, t, hh] = part[lanes, t ^ ct, hh ^ ch]
    n_shfl = 0
    for t in range(4):
        for hh in range(R2):
            s[:, t, hh] = s[:, t, hh] + shfl_xor(s[:, t + 4, hh], 16); n_shfl += 1
    for t in range(2):
        for hh in range(R2):
            s[:, t, hh] = s[:, t, hh] + shfl_xor(s[:, t + 2, hh], 8); n_shfl += 1
    for hh in range(R2):
        s[:, 0, hh] = s[:, 0, hh] + shfl_xor(s[:, 1, hh], 4); n_shfl += 1
    N4 = R2 // 2 if HB >= 1 else 1
    for hh in range(N4):
        s[:, 0, hh] = s[:, 0, hh] + shfl_xor(s[:, 0, hh + N4 if HB >= 1 else hh], 1); n_shfl += 1
    N5 = R2 // 4 if HB >= 2 else 1
    for hh in range(N5):
        s[:, 0, hh] = s[:, 0, hh] + shfl_xor(s[:, 0, hh + N5 if HB >= 2 else hh], 2); n_shfl += 1
    assert n_shfl == 7 * R2 + N4 + N5               # R2 = 4: 31 shuffles for 32 sums (160 with full butterflies)
    seen = set()
    for lane in range(32):
        owner = HB >= 2 or ((lane & 2) == 0 if HB == 1 else (lane & 3) == 0)
        for hh in range(N5):
            head = hh ^ int(ch[lane])
            assert s[lane, 0, hh].tobytes() == ref[int(ct[lane])][head].tobytes(), (lane, hh)
            if owner:
                assert (int(ct[lane]), head) not in seen    # every (row, head) is stored by exactly one lane
                seen.add((int(ct[lane]), head))
    assert seen == {(r, h) for r in range(8) for h in range(R2)}


def rs_stage(a: np.ndarray, up: np.ndarray, mask: int) -> np.ndarray:
    """ps_rs_stage<N>: a[lane][N] -> [lane][N / 2]; lanes with `up` keep the upper half."""
    n = a.shape[1] // 2
    keep = np.where(up[:, None], a[:, n:], a[:, :n])
    send = np.where(up[:, None], a[:, :n], a[:, n:])
    return (keep + send[np.arange(32) ^ mask]).astype(np.float32)


def test_pv_tile_sixty_four_sums():
    """ps_k_attn_pv_tile: 64 sums (flat index qi + 8 * di) -> 2 per lane; lane l ends with di = l >> 2 and
    qi = 4 * (l & 1) + (l & 2) + {0, 1}."""
    rng = np.random.default_rng(3)
    part = rand_parts(rng, 32, 64)
    ref = [butterfly(part[:, k]) for k in range(64)]
    lanes = np.arange(32)
    a = part.copy()
    a = rs_stage(a, (lanes & 16) != 0, 16)
    a = rs_stage(a, (lanes & 8) != 0, 8)
    a = rs_stage(a, (lanes & 4) != 0, 4)
    a = rs_stage(a, (lanes & 1) != 0, 1)
    a = rs_stage(a, (lanes & 2) != 0, 2)
    assert a.shape == (32, 2)
    seen = set()
    for lane in range(32):
        di, qb = lane >> 2, 4 * (lane & 1) + (lane & 2)
        for e in range(2):
            k = (qb + e) + 8 * di
            assert a[lane, e].tobytes() == ref[k].tobytes(), (lane, e)
            seen.add(k)
    assert seen == set(range(64))
