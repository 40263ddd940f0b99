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
    # digit products through FP64 BLAS: every partial sum is an integer below 2^53, so they are exact
    FA, FB = PA.astype(np.float64), PB.astype(np.float64)
    D = [np.zeros((A.shape[0], B.shape[0])) for _ in range(S)]
    for p in range(S):
        for q in range(S - p):
            D[p + q] += FA[p] @ FB[q].T
    for u in range(S):
        assert np.max(np.abs(D[u])) < 2**31, "a diagonal overflows the s32 accumulator"
        D[u] = D[u].astype(np.int64)
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


def test_digit_count_of_the_factorisation():
    """Why potrf / trtri run with 7 digits: the fused potrf + trtri recursion of exq_api.cu (factor_rec), restated
    in numpy with the int8 arithmetic for inner dimensions >= 512, on the benchmark's covariance (N = 2048 here).
    56 bits reproduce the FP64 recursion's log-likelihood to 1e-13 and K^-1 y to a few 1e-13; with 48 bits in the
    two triangular-inverse merges K^-1 y moves by ~1e-10 -- still inside BASELINE.json's 1e-8, but 100 x closer."""
    import scipy.linalg as sla

    from oracle import gp_oracle as orc
    import oracle.mogp_emulator as mogp

    def mm(A, Bt, S):
        if S and A.shape[1] >= 512:
            return product(np.ascontiguousarray(A), np.ascontiguousarray(Bt), S)
        return A @ Bt.T

    def factor(A, Li, r0, n, sp, st):
        if n == 128:
            L = np.linalg.cholesky(A[r0:r0 + n, r0:r0 + n])
            A[r0:r0 + n, r0:r0 + n] = L
            Li[r0:r0 + n, r0:r0 + n] = sla.solve_triangular(L, np.eye(n), lower=True)
            return
        n1 = n // 2
        a, b, c = r0, r0 + n1, r0 + n
        factor(A, Li, a, n1, sp, st)
        L21 = mm(A[b:c, a:b], Li[a:b, a:b], sp)             # (1) A21 Linv11^T
        A[b:c, b:c] -= mm(L21, L21, sp)                     # (2) SYRK
        T = mm(L21, Li[a:b, a:b].T, st)                     # (3) L21 Linv11
        factor(A, Li, b, n - n1, sp, st)
        Li[b:c, a:b] = -mm(Li[b:c, b:c], T.T, st)           # (4) -Linv22 T
        A[b:c, a:b] = L21

    N, d = 2048, 8
    X, y = orc.synthetic_problem(N, d)
    K = 1.7 * mogp.Kernel.Matern52().kernel_f(X, X, np.full(d, -2 * np.log(0.5)))
    out = {}
    for name, sp, st in (("fp64", 0, 0), ("7/7", 7, 7), ("7/6", 7, 6)):
        A, Li = K.copy(), np.zeros_like(K)
        factor(A, Li, 0, N, sp, st)
        t = Li @ y
        nll = 0.5 * (2 * np.sum(np.log(np.diag(A))) + t @ t + N * np.log(2 * np.pi))
        out[name] = (nll, Li.T @ t)
    ref_nll, ref_alpha = out["fp64"]
    err = {k: (abs(v[0] - ref_nll) / abs(ref_nll), np.max(np.abs(v[1] - ref_alpha)) / np.max(np.abs(ref_alpha)))
           for k, v in out.items() if k != "fp64"}
    assert err["7/7"][0] < 1e-13 and err["7/7"][1] < 5e-12, err
    assert err["7/6"][0] < 1e-9 and err["7/6"][1] < 1e-8, err
    assert err["7/6"][1] > 10 * err["7/7"][1], err
