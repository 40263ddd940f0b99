// nugget = "pivot": the pivoted Cholesky factorisation mogp-emulator reaches through
// scipy.linalg.lapack.dpstrf (mogp linalg.cholesky.pivot_cholesky, behind GaussianProcess.fit when
// nugget_type == "pivot"; exauq/core/emulators.py:80-82, 129-137 pass the option through, tests/unit/test_emulators.py
// :35, 683, 706 exercise it).  Restated on the device as LAPACK's unblocked lower variant (dpstf2):
//
//   dots_i = sum_{k<j} L_ik^2 ;  pivot = FIRST arg max_{i>=j} (a_ii - dots_i) ;  stop when that maximum is
//   <= n * eps * max_i a_ii  (eps = 2^-53, DLAMCH('Epsilon')): rank = j ;  symmetric interchange j <-> pivot ;
//   L_jj = sqrt(.) ;  L_ij = (a_ij - sum_{k<j} L_ik L_jk) * (1 / L_jj)   for i > j   (left-looking DGEMV + DSCAL).
//
// One cooperative launch (grid-wide barriers between the three phases of a step) -- this is the robustness option of
// the interface, not the hot path: O(n^3 / 3) memory-bound work, ~0.1 s at n = 4096.  The same kernel with
// pivoting = 0 factors the leading `stop_at` columns of an already permuted matrix, after which
// pivot_fill_kernel applies mogp's rule for the rank-deficient remainder (diagonal entries
// L[rank-1][rank-1] / ((rank+1)(rank+2)...(i+1)), so that log det stays defined) and trtri_lower_kernel forms the
// inverse factor the rest of the library works with.
#pragma once
#include <cooperative_groups.h>
#include "exq_common.cuh"

namespace exq {
namespace cg = cooperative_groups;

struct PivArgs {
    double* A;        // [n][ld] lower triangle (row-major); overwritten by L
    int64_t ld;
    int n;
    int pivoting;     // 1: dpstf2 with diagonal pivoting; 0: no interchanges, stop after stop_at columns
    int stop_at;
    double* dots;     // [n]
    double* work;     // [n]
    int* piv;         // [n] out: row / column i of the factored matrix is row / column piv[i] of the input
    int* rank;        // out
    double* cand_v;   // [gridDim.x] per-block arg max
    int* cand_i;      // [gridDim.x]
};

constexpr int PIV_THREADS = 256;

__global__ void __launch_bounds__(PIV_THREADS) pivot_chol_kernel(PivArgs p) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double piv_rowj[];            // row j of L (columns < j)
    __shared__ double red_v[PIV_THREADS / 32];
    __shared__ int red_i[PIV_THREADS / 32];
    __shared__ double bc_v;
    __shared__ int bc_i;
    const int n = p.n;
    const int64_t ld = p.ld;
    double* A = p.A;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gtid = blockIdx.x * blockDim.x + tid, gsize = gridDim.x * blockDim.x;
    const int gwarp = gtid >> 5, nwarps = gsize >> 5;

    // first-index arg max of (v, i) pairs over the block, NaN never wins; result broadcast through bc_v / bc_i
    auto block_argmax = [&](double v, int i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        if (lane == 0) { red_v[warp] = v; red_i[warp] = i; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < PIV_THREADS / 32; w++)
                if (red_v[w] > v || (red_v[w] == v && red_i[w] < i)) { v = red_v[w]; i = red_i[w]; }
            bc_v = v; bc_i = i;
        }
        __syncthreads();
    };

    for (int i = gtid; i < n; i += gsize) { p.dots[i] = 0.0; p.piv[i] = i; }
    // stopping threshold from the largest diagonal entry of the input (every block computes it for itself)
    {
        double v = -INFINITY; int bi = 0x7fffffff;
        for (int i = tid; i < n; i += PIV_THREADS) {
            const double a = A[(int64_t)i * ld + i];
            if (a > v) { v = a; bi = i; }
        }
        block_argmax(v, bi);
    }
    const double dmax = bc_v;
    const double dstop = (double)n * 1.1102230246251565e-16 * dmax;
    int rank = n;
    if (p.pivoting && !(dmax > 0.0)) rank = 0;       // dpstf2: INFO = 1, RANK = 0
    grid.sync();

    for (int j = 0; j < n && rank == n; j++) {
        if (!p.pivoting && j >= p.stop_at) { rank = j; break; }
        // ---- phase A: running sums of squares, candidate pivots, arg max ----------------------------------
        {
            double v = -INFINITY; int bi = 0x7fffffff;
            for (int i = j + gtid; i < n; i += gsize) {
                double dt = p.dots[i];
                if (j > 0) {
                    const double a = A[(int64_t)i * ld + j - 1];
                    dt = __dadd_rn(dt, __dmul_rn(a, a));
                    p.dots[i] = dt;
                }
                const double w = A[(int64_t)i * ld + i] - dt;
                p.work[i] = w;
                if (w > v) { v = w; bi = i; }
            }
            block_argmax(v, bi);
            if (tid == 0) { p.cand_v[blockIdx.x] = bc_v; p.cand_i[blockIdx.x] = bc_i; }
        }
        grid.sync();
        int pvt = j;
        double ajj;
        if (p.pivoting) {
            double v = -INFINITY; int bi = 0x7fffffff;
            for (int b = tid; b < (int)gridDim.x; b += PIV_THREADS) {
                const double cv = p.cand_v[b]; const int ci = p.cand_i[b];
                if (cv > v || (cv == v && ci < bi)) { v = cv; bi = ci; }
            }
            block_argmax(v, bi);
            ajj = bc_v; pvt = bc_i;
            if (!(ajj > dstop)) { rank = j; break; }          // also taken for NaN (DISNAN in dpstf2)
        } else {
            ajj = p.work[j];
            if (!(ajj > 0.0)) { rank = j; break; }
        }
        // ---- phase B: symmetric interchange j <-> pvt in the lower triangle ---------------------------------
        if (pvt != j) {
            // the ranges: k < j (row segments), j < i < pvt (column j against row pvt), i > pvt (columns j and pvt)
            for (int e = gtid; e < n; e += gsize) {
                if (e < j) {
                    const double t = A[(int64_t)j * ld + e];
                    A[(int64_t)j * ld + e] = A[(int64_t)pvt * ld + e];
                    A[(int64_t)pvt * ld + e] = t;
                } else if (e > j && e < pvt) {
                    const double t = A[(int64_t)e * ld + j];
                    A[(int64_t)e * ld + j] = A[(int64_t)pvt * ld + e];
                    A[(int64_t)pvt * ld + e] = t;
                } else if (e > pvt) {
                    const double t = A[(int64_t)e * ld + j];
                    A[(int64_t)e * ld + j] = A[(int64_t)e * ld + pvt];
                    A[(int64_t)e * ld + pvt] = t;
                } else if (e == j) {
                    const double t = A[(int64_t)j * ld + j];
                    A[(int64_t)j * ld + j] = A[(int64_t)pvt * ld + pvt];
                    A[(int64_t)pvt * ld + pvt] = t;
                    const double td = p.dots[j]; p.dots[j] = p.dots[pvt]; p.dots[pvt] = td;
                    const int tp = p.piv[j]; p.piv[j] = p.piv[pvt]; p.piv[pvt] = tp;
                }
            }
            grid.sync();
        }
        // ---- phase C: column j --------------------------------------------------------------------------------
        const double sq = sqrt(ajj);
        const double inv = 1.0 / sq;
        for (int k = tid; k < j; k += PIV_THREADS) piv_rowj[k] = A[(int64_t)j * ld + k];
        __syncthreads();
        if (gtid == 0) A[(int64_t)j * ld + j] = sq;
        for (int i = j + 1 + gwarp; i < n; i += nwarps) {
            const double* row = A + (int64_t)i * ld;
            double s = 0.0;
            for (int k = lane; k < j; k += 32) s = fma(row[k], piv_rowj[k], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) A[(int64_t)i * ld + j] = (row[j] - s) * inv;
        }
        grid.sync();
    }
    if (gtid == 0) *p.rank = rank;
}

// mogp pivot_cholesky after dpstrf: L = tril(L); L[i][i] = L[rank-1][rank-1] / prod_{m=rank+1}^{i+1} m  for i >= rank.
// Also zeroes the strictly upper part of the leading n x n block (the assembly wrote covariance entries there).
__global__ void pivot_fill_kernel(double* __restrict__ A, int64_t ld, int n, const int* __restrict__ rank_dev) {
    const int rank = *rank_dev;
    const int i = blockIdx.x;
    if (i >= n) return;
    for (int c = i + 1 + threadIdx.x; c < n; c += blockDim.x) A[(int64_t)i * ld + c] = 0.0;
    if (threadIdx.x == 0 && i >= rank && rank >= 1) {
        double div = 1.0;
        for (int m = rank + 1; m <= i + 1; m++) div *= (double)m;
        A[(int64_t)i * ld + i] = A[(int64_t)(rank - 1) * ld + rank - 1] / div;
    }
}

// Linv = L^-1 for a general lower-triangular L (n x n inside an [Np][ld] buffer whose padding rows are identity rows):
// one block per column, forward substitution with the column kept in shared memory.  Linv must be zeroed beforehand.
__global__ void __launch_bounds__(128) trtri_lower_kernel(const double* __restrict__ L, double* __restrict__ Linv, int64_t ld,
                                                          int n, int np) {
    extern __shared__ double tri_x[];               // x[c..n)
    __shared__ double part[4];
    const int c = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (c >= n) {                                    // padding: identity
        if (c < np && tid == 0) Linv[(int64_t)c * ld + c] = 1.0;
        return;
    }
    for (int i = c; i < n; i++) {
        const double* row = L + (int64_t)i * ld;
        double s = 0.0;
        for (int k = c + tid; k < i; k += 128) s = fma(row[k], tri_x[k - c], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) part[warp] = s;
        __syncthreads();
        if (tid == 0) {
            const double tot = (part[0] + part[1]) + (part[2] + part[3]);
            const double x = ((i == c ? 1.0 : 0.0) - tot) / row[i];
            tri_x[i - c] = x;
            Linv[(int64_t)i * ld + c] = x;
        }
        __syncthreads();
    }
}

}  // namespace exq
