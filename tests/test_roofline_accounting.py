"""The int8 operation count behind ``roofline.achieved`` is re-derivable: the library's own count
(``exq_int8_gemm_ops``, the tile walk of csrc/ozaki_gemm.cuh: ``decode_gemm_item`` / ``tile_krange``)
is compared with an independent restatement of that walk, for every product of the potrf + trtri
recursion and LAUUM at the benchmark size.  Host arithmetic only -- CPU suite.

Round 1 counted LAUUM (lower tiles, triangular k range) as 2/3 * 1/2 = 1/3 of the cube where the
kernel issues 1/6: the printed roofline fraction (0.93) was 0.72-0.77 in fact.
"""

import numpy as np
import pytest

from exauq_toolbox_b200 import _lib

BLK_C, BLK_R, BLK_K = 128, 64, 64        # output tile rows / columns, k bytes per counted block (ozaki_i8.cuh)
PIPE_K = 128                             # the kernels walk k in blocks of 128 bytes: ranges widen to whole blocks
KR_FULL, KR_LE_J, KR_GE_J, KR_LE_I, KR_GE_I = range(5)


def kblocks(M, N, K, krange, lower_only):
    """k blocks walked for one system -- written from the description in DESIGN.md section 4a, not from the C++."""
    nI, nJ, nkb = M // BLK_C, N // BLK_R, K // BLK_K
    total = 0
    for i in range(nI):
        rows = (i * BLK_C, i * BLK_C + BLK_C - 1)
        for j in range(nJ):
            cols = (j * BLK_R, j * BLK_R + BLK_R - 1)
            if lower_only and cols[0] > rows[1]:
                continue                                    # tile strictly above the diagonal
            lo, hi = 0, nkb
            if krange == KR_LE_J:                           # B lower triangular: k <= column
                hi = min(nkb, cols[1] // BLK_K + 1)
            elif krange == KR_GE_J:                         # k >= column
                lo = cols[0] // BLK_K
            elif krange == KR_LE_I:                         # A lower triangular: k <= row
                hi = min(nkb, rows[1] // BLK_K + 1)
            elif krange == KR_GE_I:
                lo = rows[0] // BLK_K
            if hi > lo:
                # first and last 128-byte block that hold a kept 64-byte block, counted in 64-byte units
                first, last = (lo * BLK_K) // PIPE_K, ((hi - 1) * BLK_K) // PIPE_K
                total += (last - first + 1) * (PIPE_K // BLK_K)
    return total


def ops(M, N, K, B, krange, lower_only, S):
    return 2.0 * BLK_C * BLK_R * BLK_K * kblocks(M, N, K, krange, lower_only) * B * (S * (S + 1) // 2)


# every int8 product of one evaluation at N = 4096 (nodes n = 2048 and n = 4096, then LAUUM)
PRODUCTS = [
    ("L21 = A21 Linv11^T", 1024, 1024, 1024, KR_LE_J, 0, 7),
    ("A22 -= L21 L21^T", 1024, 1024, 1024, KR_FULL, 1, 7),
    ("T = L21 Linv11", 1024, 1024, 1024, KR_GE_J, 0, 7),
    ("Linv21 = -Linv22 T", 1024, 1024, 1024, KR_LE_I, 0, 7),
    ("L21 (top)", 2048, 2048, 2048, KR_LE_J, 0, 7),
    ("SYRK (top)", 2048, 2048, 2048, KR_FULL, 1, 7),
    ("T (top)", 2048, 2048, 2048, KR_GE_J, 0, 7),
    ("Linv21 (top)", 2048, 2048, 2048, KR_LE_I, 0, 7),
    ("LAUUM", 4096, 4096, 4096, KR_GE_I, 1, 6),
    ("odd split", 1024, 1152, 1152, KR_LE_J, 0, 7),
]


@pytest.mark.parametrize("name,M,N,K,kr,lower,S", PRODUCTS)
def test_library_count_equals_the_tile_walk(name, M, N, K, kr, lower, S):
    for B in (1, 15):
        got = _lib.int8_gemm_ops(M, N, K, B, kr, bool(lower), S)
        assert got == ops(M, N, K, B, kr, lower, S), name


def test_lauum_issues_a_sixth_of_the_cube_not_a_third():
    n, S = 4096, 6
    issued = _lib.int8_gemm_ops(n, n, n, 1, KR_GE_I, True, S)
    cube = 2.0 * n ** 3 * (S * (S + 1) // 2)
    frac = issued / cube
    assert 1 / 6 < frac < 1 / 6 * 1.12, frac            # 1/6 plus the overhang of the 128 x 64 diagonal tiles (9.6 % at n = 4096)
    assert abs(frac - 1 / 3) > 0.1


def test_triangular_and_syrk_products_issue_about_half():
    n, S = 2048, 7
    cube = 2.0 * n ** 3 * (S * (S + 1) // 2)
    for kr, lower in ((KR_LE_J, 0), (KR_GE_J, 0), (KR_LE_I, 0), (KR_FULL, 1)):
        frac = _lib.int8_gemm_ops(n, n, n, 1, kr, bool(lower), S) / cube
        assert 0.5 <= frac < 0.56, (kr, lower, frac)


def test_bad_shapes_are_refused():
    assert _lib.int8_gemm_ops(100, 64, 64, 1, 0, False, 7) < 0
    assert _lib.int8_gemm_ops(128, 64, 64, 1, 9, False, 7) < 0


def test_whole_recursion_matches_the_bench_accounting():
    """Issued int8 operations of one evaluation at N = 4096 (what bench.py divides by the kernel time), and the
    FP64-equivalent flop count beside it: potrf + trtri of the two int8 levels + LAUUM."""
    total = 0.0
    for name, M, N, K, kr, lower, S in PRODUCTS[:4]:
        total += 2 * ops(M, N, K, 1, kr, lower, S)         # two nodes of size 2048
    for name, M, N, K, kr, lower, S in PRODUCTS[4:9]:
        total += ops(M, N, K, 1, kr, lower, S)
    lib_total = (2 * sum(_lib.int8_gemm_ops(M, N, K, 1, kr, bool(lo), S) for _, M, N, K, kr, lo, S in PRODUCTS[:4])
                 + sum(_lib.int8_gemm_ops(M, N, K, 1, kr, bool(lo), S) for _, M, N, K, kr, lo, S in PRODUCTS[4:9]))
    assert total == lib_total
    # algorithmic FP64 flops of the same products: 4 x n^3 per node (n = child size) + N^3 / 3
    fp64 = 2 * 4 * 1024.0 ** 3 + 4 * 2048.0 ** 3 + 4096.0 ** 3 / 3
    ratio = total / fp64
    # 28 (or 21) int8 products per FP64 product, times the tile overhang of the triangular ranges (< 8 %)
    assert 24.0 < ratio < 28.0 * 1.08, ratio
    assert np.isfinite(ratio)


# ---- balanced work lists (csrc/ozaki_gemm.cuh: make_schedule) ------------------------------------------------------------
SCHED_NONE = 0xFFFFFFFF


def _tile_weight(M, N, K, kr, I, J):
    """128-byte k blocks the kernel walks for output tile (I, J) -- the same widening as kblocks() above."""
    nkb = K // BLK_K
    rows = (I * BLK_C, I * BLK_C + BLK_C - 1)
    cols = (J * BLK_R, J * BLK_R + BLK_R - 1)
    lo, hi = 0, nkb
    if kr == KR_LE_J:
        hi = min(nkb, cols[1] // BLK_K + 1)
    elif kr == KR_GE_J:
        lo = cols[0] // BLK_K
    elif kr == KR_LE_I:
        hi = min(nkb, rows[1] // BLK_K + 1)
    elif kr == KR_GE_I:
        lo = rows[0] // BLK_K
    return ((hi - 1) * BLK_K) // PIPE_K - (lo * BLK_K) // PIPE_K + 1 if hi > lo else 0


def _loads(sched, grid, M, N, K, kr, B=None):
    """k blocks per CTA when CTA c walks entries c, c + grid, ... (entries of systems >= B are skipped)"""
    load = np.zeros(grid)
    for i, it in enumerate(sched):
        if it == SCHED_NONE:
            continue
        b, I, J = int(it) >> 20, (int(it) >> 10) & 1023, int(it) & 1023
        if B is not None and b >= B:
            continue
        load[i % grid] += _tile_weight(M, N, K, kr, I, J)
    return load


@pytest.mark.parametrize("name,M,N,K,kr,lower,S", PRODUCTS)
def test_work_list_covers_every_kept_tile_once_and_is_balanced(name, M, N, K, kr, lower, S):
    B, grid = 15, 148
    sched = _lib.int8_gemm_schedule(M, N, K, B, kr, bool(lower), grid)
    assert len(sched) % grid == 0
    items = sched[sched != SCHED_NONE]
    want = {(b << 20) | (i << 10) | j for b in range(B) for i in range(M // BLK_C) for j in range(N // BLK_R)
            if not (lower and j * BLK_R > i * BLK_C + BLK_C - 1)}
    assert len(items) == len(want) and set(int(x) for x in items) == want, name
    # the same number of k blocks as the arithmetic walk issues
    load = _loads(sched, grid, M, N, K, kr)
    assert load.sum() * BLK_C * BLK_R * PIPE_K * 2.0 * (S * (S + 1) // 2) == _lib.int8_gemm_ops(M, N, K, B, kr, bool(lower), S)
    # systems stay in sequence (their digit planes share L2): per CTA the system index never decreases
    for c in range(grid):
        mine = sched[c::grid]
        sysid = (mine[mine != SCHED_NONE] >> 20).astype(int)
        assert np.all(np.diff(sysid) >= 0), name
    # balance: the busiest CTA within one heavy tile of the average (the arithmetic walk: 7-38 % above it)
    heaviest = max(_tile_weight(M, N, K, kr, i, j) for i in range(M // BLK_C) for j in range(N // BLK_R))
    assert load.max() - load.mean() <= heaviest, (name, load.max(), load.mean())
    if len(items) >= 20 * grid:
        assert load.max() <= 1.03 * load.mean(), (name, load.max() / load.mean())


def test_prefix_of_a_work_list_serves_a_smaller_batch():
    """The library builds a list for the next power of two of the batch and uses the prefix of the systems it has."""
    M = N = K = 2048
    grid = 148
    for kr, lower in ((KR_LE_J, 0), (KR_FULL, 1), (KR_GE_I, 1)):
        full = _lib.int8_gemm_schedule(M, N, K, 16, kr, bool(lower), grid)
        kept = int(np.count_nonzero(full != SCHED_NONE)) // 16
        for B in (9, 13, 15):
            n = -(-B * kept // grid) * grid                      # rounds that hold the first B systems
            prefix = full[:n]
            got = prefix[(prefix != SCHED_NONE) & ((prefix >> 20) < B)]
            own = _lib.int8_gemm_schedule(M, N, K, B, kr, bool(lower), grid)
            assert set(int(x) for x in got) == set(int(x) for x in own[own != SCHED_NONE])
            load = _loads(prefix, grid, M, N, K, kr, B)
            assert load.max() <= 1.05 * load.mean(), (kr, B, load.max() / load.mean())


def test_work_list_refuses_bad_shapes():
    with pytest.raises(ValueError):
        _lib.int8_gemm_schedule(100, 64, 128, 1, 0, False, 148)
    with pytest.raises(ValueError):
        _lib.int8_gemm_schedule(128, 64, 64, 1, 0, False, 148)     # k blocks are 128 bytes
