"""The arithmetic of the int8 tensor-core kernels (csrc/ozaki_i8.cuh, csrc/ozaki_gemm.cuh), restated in numpy
so that it can be checked without a GPU: balanced base-256 digits through the bias trick, exact integer
diagonals D_t, recombination in two int64 groups, one rounding to FP64.  The kernels are compared with the
FP64 DMMA path and the oracle on the GPU (tests/test_gpu_int8.py); this file pins the scheme itself: the
digits reproduce the scaled value exactly, every diagonal fits an int32, and the error of a product is
bounded by ~K 2^-8S max|a_row| max|b_row|."""

import numpy as np
import pytest


def scale_exponent(mx):
    """oz::scale_exponent: smallest e (plus the 0.49 guard) with |v| 2^-e < 0.49."""
    e = np.where(mx > 0, np.floor(np.log2(np.where(mx > 0, mx, 1.0))) + 2, 0.0)
    e = np.where((mx > 0) & (mx * np.exp2(-e) >= 0.49), e + 1, e)
    return np.clip(e, -900, 900)


def digits(A, S):
    """oz::digits_of per element: x = rint(v 2^(8S - e)) + 0x80..80, byte (S-1-p) ^ 0x80 = digit of plane p."""
    e = scale_exponent(np.max(np.abs(A), axis=1))
    x = np.rint(np.ldexp(A * np.exp2(-e)[:, None], 8 * S)).astype(np.int64)
    bias = int("80" * S, 16)
    u = (x + bias).astype(np.uint64)
    planes = []
    for p in range(S):
        byte = ((u >> np.uint64(8 * (S - 1 - p))) & np.uint64(0xFF)).astype(np.int64) ^ 0x80
        planes.append(np.where(byte >= 128, byte - 256, byte))          # two's complement int8
    return np.array(planes), e, x


def product(A, B, S):
    """A @ B.T as the kernels compute it."""
    PA, eA, _ = digits(A, S)
    PB, eB, _ = digits(B, S)
    D = [np.zeros((A.shape[0], B.shape[0]), dtype=np.int64) for _ in range(S)]
    for p in range(S):
        for q in range(S - p):
            D[p + q] += PA[p] @ PB[q].T
    for d in D:
        assert np.max(np.abs(d)) < 2**31, "a diagonal overflows the s32 accumulator"
    hi = np.zeros_like(D[0])
    lo = np.zeros_like(D[0])
    for u in range(min(4, S)):
        hi = hi * 256 + D[u]
    for u in range(4, S):
        lo = lo * 256 + D[u]
    z = hi.astype(np.float64) + lo.astype(np.float64) * 2.0 ** (-8 * (S - 4))
    return z * np.exp2(eA - 40)[:, None] * np.exp2(eB)[None, :]


@pytest.mark.parametrize("S", [5, 6, 7])
def test_digits_are_exact_and_in_range(S):
    rng = np.random.default_rng(S)
    A = rng.normal(size=(37, 200)) * 10.0 ** rng.uniform(-6, 6, (37, 1))
    A[3] = 0.0                                   # an all-zero row
    A[5, 7] = A[5].max() * 1.0000001             # a value right at the top of its row
    P, e, x = digits(A, S)
    assert P.min() >= -128 and P.max() <= 127
    recon = sum(P[p].astype(np.int64) * 256 ** (S - 1 - p) for p in range(S))
    assert np.array_equal(recon, x)              # the digits are the scaled, rounded value -- exactly
    assert np.max(np.abs(A - np.ldexp(x.astype(np.float64), -8 * S) * np.exp2(e)[:, None]) /
                  np.maximum(np.max(np.abs(A), axis=1, keepdims=True), 1e-300)) <= 2.0 ** (-8 * S + 2)   # half a unit of the last digit over a scaled row maximum >= 1/8


@pytest.mark.parametrize("S,K", [(5, 512), (6, 4096), (7, 4096)])
def test_product_error_bound(S, K):
    rng = np.random.default_rng(100 + S)
    A = rng.normal(size=(24, K)) * 10.0 ** rng.uniform(-2, 2, (24, K))
    B = np.abs(rng.normal(size=(16, K)))
    want = (A.astype(np.longdouble) @ B.T.astype(np.longdouble)).astype(np.float64)
    got = product(A, B, S)
    bound = K * 2.0 ** (-8 * S + 2) * np.max(np.abs(A), axis=1)[:, None] * np.max(np.abs(B), axis=1)[None, :]
    assert np.all(np.abs(got - want) <= bound + 4 * np.finfo(float).eps * np.abs(want))


def test_worst_case_diagonal_fits_s32_up_to_k_16384():
    # all digits at their extreme value: |D_t| <= (t pairs) * 128 * 128 * K
    for S in (6, 7):
        assert S * 128 * 128 * 16384 < 2**31
