// FP64 tensor-core (DMMA.8x8x4) tile GEMM used by every O(N^3) step of the engine:
// TRSM-by-inverse, SYRK trailing update, triangular inverse merge, LAUUM (K^-1 = Linv^T Linv)
// and the predictive-variance product Z = Linv * Kc^T with a fused column sum-of-squares
// epilogue.  There is no reference counterpart: the reference reaches LAPACK dpotrf /
// cho_solve through mogp ("BLAS/LAPACK library call sites").
//
//   C[i, j] (op)= sum_k A(i, k) * B(k, j)      i < M, j < N, k in a per-tile range
//
// CTA tile 128 x 128, k-tile 16, 8 warps as 2 (m) x 4 (n), warp tile 64 x 32 held as
// 8 x 4 DMMA accumulators (64 doubles / thread).  Operand tiles are staged with a
// 4-stage cp.async (LDGSTS) ring in padded shared memory so that all fragment loads are
// bank-conflict free.  All extents are multiples of the tile; no bounds checks.
#pragma once
#include "exq_common.cuh"

namespace exq {

enum : int { LAYOUT_K = 0, LAYOUT_MN = 1 };          // which index is contiguous in memory
enum : int { KR_FULL = 0, KR_LE_J = 1, KR_GE_J = 2, KR_LE_I = 3, KR_GE_I = 4 };
enum : int { C_STORE = 0, C_SUB = 1, C_NEG = 2, C_COLSUMSQ = 3 };

struct GemmArgs {
    const double* A;   // LAYOUT_K: A[i*lda + k]    LAYOUT_MN: A[k*lda + i]
    const double* B;   // LAYOUT_K: B[j*ldb + k]    LAYOUT_MN: B[k*ldb + j]
    double* C;         // C[i*ldc + j]; for C_COLSUMSQ: partial[(i/128)*ldc + j]
    int64_t lda, ldb, ldc;
    int64_t strideA, strideB, strideC;   // batch strides (elements), blockIdx.z
    int M, N, K;       // multiples of 128 / 128 / 16
    int krange;        // KR_*: restricts the k loop per output tile (triangular operands)
    int lower_only;    // skip tiles strictly above the diagonal
    int cmode;         // C_*
    int raster;        // 0: blockIdx.x -> column tile, blockIdx.y -> row tile.
                       // G > 0: 1-D grid over (row tile, column tile); column tiles are taken in
                       //    groups of G and, inside a group, row tiles heaviest-first.  The G
                       //    B-operand tiles of a group plus the whole A operand stay in L2, so
                       //    HBM traffic drops to about the algorithmic bytes.
};

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 16, GEMM_STAGES = 4, GEMM_THREADS = 256;
constexpr int GEMM_LDS_K = GEMM_BK + 4;     // padded row stride of a K-contiguous tile
constexpr int GEMM_LDS_MN = GEMM_BM + 4;    // padded row stride of an M/N-contiguous tile
constexpr int GEMM_TILE_ELEMS = GEMM_BM * GEMM_LDS_K;   // 2560 >= 16*132 = 2112
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * 2 * GEMM_TILE_ELEMS * (int)sizeof(double);

template <int LAYOUT>
__device__ __forceinline__ void gemm_load_tile(double* s, const double* g, int64_t ld, int row0,
                                               int k0, int tid) {
    if (LAYOUT == LAYOUT_K) {
        // 128 rows x 16 k, k contiguous: 1024 16-byte chunks
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int c = tid + i * GEMM_THREADS;
            int r = c >> 3, kc = c & 7;
            cp_async16(s + r * GEMM_LDS_K + 2 * kc, g + (int64_t)(row0 + r) * ld + k0 + 2 * kc);
        }
    } else {
        // 16 k-rows x 128 m, m contiguous
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int c = tid + i * GEMM_THREADS;
            int kr = c >> 6, mc = c & 63;
            cp_async16(s + kr * GEMM_LDS_MN + 2 * mc, g + (int64_t)(k0 + kr) * ld + row0 + 2 * mc);
        }
    }
}

template <int AL, int BL>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_dmma_kernel(GemmArgs p) {
    extern __shared__ __align__(16) double gemm_smem[];
    // heavy tiles first: with triangular k ranges the work grows with it (or shrinks with jt)
    int it, jt;
    if (p.raster > 0) {
        const int nI = p.M / GEMM_BM, nJ = p.N / GEMM_BN, G = p.raster;
        const int b = blockIdx.x, per_group = G * nI;
        const int grp = b / per_group, r = b - grp * per_group;
        const int gw = min(G, nJ - grp * G);
        const int ir = r / gw;
        jt = grp * G + (r - ir * gw);
        it = (p.krange == KR_LE_I) ? (nI - 1 - ir) : ir;
    } else {
        it = (p.krange == KR_LE_I) ? (gridDim.y - 1 - blockIdx.y) : blockIdx.y;
        jt = (p.krange == KR_LE_J) ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
    }
    if (p.lower_only && jt > it) return;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;   // 2 x 4 warps

    const double* A = p.A + (int64_t)blockIdx.z * p.strideA;
    const double* B = p.B + (int64_t)blockIdx.z * p.strideB;
    double* C = p.C + (int64_t)blockIdx.z * p.strideC;

    int kt_lo = 0, kt_hi = p.K / GEMM_BK;
    const int tiles_per_blk = GEMM_BM / GEMM_BK;   // 8 k-tiles per 128 block
    if (p.krange == KR_LE_J) kt_hi = min(kt_hi, (jt + 1) * tiles_per_blk);
    else if (p.krange == KR_GE_J) kt_lo = jt * tiles_per_blk;
    else if (p.krange == KR_LE_I) kt_hi = min(kt_hi, (it + 1) * tiles_per_blk);
    else if (p.krange == KR_GE_I) kt_lo = it * tiles_per_blk;
    const int nkt = kt_hi - kt_lo;

    double* sA = gemm_smem;
    double* sB = gemm_smem + GEMM_STAGES * GEMM_TILE_ELEMS;
    const int m0 = it * GEMM_BM, n0 = jt * GEMM_BN;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue
#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; s++) {
        if (s < nkt) {
            gemm_load_tile<AL>(sA + s * GEMM_TILE_ELEMS, A, p.lda, m0, (kt_lo + s) * GEMM_BK, tid);
            gemm_load_tile<BL>(sB + s * GEMM_TILE_ELEMS, B, p.ldb, n0, (kt_lo + s) * GEMM_BK, tid);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < nkt; kt++) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + GEMM_STAGES - 1;
            if (nk < nkt) {
                int s = nk % GEMM_STAGES;
                gemm_load_tile<AL>(sA + s * GEMM_TILE_ELEMS, A, p.lda, m0, (kt_lo + nk) * GEMM_BK, tid);
                gemm_load_tile<BL>(sB + s * GEMM_TILE_ELEMS, B, p.ldb, n0, (kt_lo + nk) * GEMM_BK, tid);
            }
            cp_async_commit();
        }
        const double* a_s = sA + (kt % GEMM_STAGES) * GEMM_TILE_ELEMS;
        const double* b_s = sB + (kt % GEMM_STAGES) * GEMM_TILE_ELEMS;
#pragma unroll
        for (int kk = 0; kk < GEMM_BK / 4; kk++) {
            double af[8], bf[4];
#pragma unroll
            for (int mf = 0; mf < 8; mf++) {
                int m = wm * 64 + mf * 8 + g, k = kk * 4 + t;
                af[mf] = (AL == LAYOUT_K) ? a_s[m * GEMM_LDS_K + k] : a_s[k * GEMM_LDS_MN + m];
            }
#pragma unroll
            for (int nf = 0; nf < 4; nf++) {
                int n = wn * 32 + nf * 8 + g, k = kk * 4 + t;
                bf[nf] = (BL == LAYOUT_K) ? b_s[n * GEMM_LDS_K + k] : b_s[k * GEMM_LDS_MN + n];
            }
#pragma unroll
            for (int mf = 0; mf < 8; mf++)
#pragma unroll
                for (int nf = 0; nf < 4; nf++) dmma884(acc[mf][nf][0], acc[mf][nf][1], af[mf], bf[nf]);
        }
    }
    cp_async_wait<0>();

    if (p.cmode == C_COLSUMSQ) {
        // column sums of squares over the 128 rows of this tile -> partial[it][n0 + col]
        __syncthreads();
        double* red = gemm_smem;   // [2][128]
#pragma unroll
        for (int nf = 0; nf < 4; nf++) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int mf = 0; mf < 8; mf++) {
                s0 += acc[mf][nf][0] * acc[mf][nf][0];
                s1 += acc[mf][nf][1] * acc[mf][nf][1];
            }
            // reduce over g (lane bits 2..4)
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
            if (g == 0) {
                int col = wn * 32 + nf * 8 + 2 * t;
                red[wm * 128 + col] = s0;
                red[wm * 128 + col + 1] = s1;
            }
        }
        __syncthreads();
        if (tid < 128) C[(int64_t)it * p.ldc + n0 + tid] = red[tid] + red[128 + tid];
        return;
    }

#pragma unroll
    for (int mf = 0; mf < 8; mf++) {
        int row = m0 + wm * 64 + mf * 8 + g;
#pragma unroll
        for (int nf = 0; nf < 4; nf++) {
            int col = n0 + wn * 32 + nf * 8 + 2 * t;
            double2* dst = reinterpret_cast<double2*>(C + (int64_t)row * p.ldc + col);
            double2 v;
            if (p.cmode == C_STORE) {
                v.x = acc[mf][nf][0]; v.y = acc[mf][nf][1];
            } else if (p.cmode == C_NEG) {
                v.x = -acc[mf][nf][0]; v.y = -acc[mf][nf][1];
            } else {
                v = *dst;
                v.x -= acc[mf][nf][0]; v.y -= acc[mf][nf][1];
            }
            *dst = v;
        }
    }
}

}  // namespace exq
