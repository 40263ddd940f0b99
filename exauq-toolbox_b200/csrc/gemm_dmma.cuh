// FP64 tensor-core (DMMA.8x8x4) tile GEMM used by every O(N^3) step of the engine:
// TRSM-by-inverse, SYRK trailing update, triangular inverse merge, LAUUM (K^-1 = Linv^T Linv)
// and the predictive-variance product Z = Linv * Kc^T with a fused column sum-of-squares
// epilogue.  There is no reference counterpart: the reference reaches LAPACK dpotrf /
// cho_solve through mogp ("BLAS/LAPACK library call sites").
//
//   C[i, j] (op)= sum_k A(i, k) * B(k, j)      i < M, j < N, k in a per-tile range
//
// CTA tile 128 x 128 (8 warps, warp tile 64 x 32) or 64 x 128 (4 warps, warp tile 32 x 64,
// 2 CTAs per SM), k-tile 16; a warp holds 32 DMMA accumulators (64 doubles / thread).  Operand tiles are staged with a
// multi-stage cp.async (LDGSTS) ring in padded shared memory so that all fragment loads are
// bank-conflict free.  All extents are multiples of the tile; no bounds checks.
#pragma once
#include <cuda.h>

#include "exq_common.cuh"

namespace exq {

enum : int { LAYOUT_K = 0, LAYOUT_MN = 1 };          // which index is contiguous in memory
enum : int { KR_FULL = 0, KR_LE_J = 1, KR_GE_J = 2, KR_LE_I = 3, KR_GE_I = 4 };
enum : int { C_STORE = 0, C_SUB = 1, C_NEG = 2, C_COLSUMSQ = 3, C_GRAD = 4 /* int8 kernel only: see ozaki_gemm.cuh */ };

struct GemmArgs {
    const double* A;   // LAYOUT_K: A[i*lda + k]    LAYOUT_MN: A[k*lda + i]
    const double* B;   // LAYOUT_K: B[j*ldb + k]    LAYOUT_MN: B[k*ldb + j]
    double* C;         // C[i*ldc + j]; for C_COLSUMSQ: partial[(i/128)*ldc + j]
    int64_t lda, ldb, ldc;
    int64_t strideA, strideB, strideC;   // batch strides (elements), blockIdx.z
    int M, N, K;       // multiples of the row tile (64 or 128) / 128 / 16
    int krange;        // KR_*: restricts the k loop per output tile (triangular operands)
    int lower_only;    // skip tiles strictly above the diagonal
    int cmode;         // C_*
    int raster;        // 0: blockIdx.x -> column tile, blockIdx.y -> row tile.
                       // G > 0: 1-D grid over (row tile, column tile); column tiles are taken in
                       //    groups of G and, inside a group, row tiles heaviest-first.  The G
                       //    B-operand tiles of a group plus the whole A operand stay in L2, so
                       //    HBM traffic drops to about the algorithmic bytes.
};

constexpr int GEMM_BN = 128, GEMM_BK = 16;
constexpr int GEMM_LDS_K = GEMM_BK + 4;      // padded row stride of a k-contiguous tile (== 4 mod 16)

// Tile configuration.  BM = 128: 8 warps (2 x 4), warp tile 64 x 32, 4 stages, 1 CTA per SM.
// BM = 64: 4 warps (2 x 2), warp tile 32 x 64, 3 stages, 2 independent CTAs per SM (the shape
// cuBLAS picks for DGEMM on this chip); also used for the small GEMMs of the recursion, where
// it quadruples the number of CTAs.
template <int BM>
struct GemmCfg {
    static constexpr int WM = 2, WN = (BM == 128) ? 4 : 2;
    static constexpr int THREADS = 32 * WM * WN;
    static constexpr int MF = BM / WM / 8, NF = GEMM_BN / WN / 8;     // DMMA tiles per warp
    static constexpr int STAGES = (BM == 128) ? 4 : 3;
    static constexpr int LDS_A_MN = BM + 4, LDS_B_MN = GEMM_BN + 4;   // m/n-contiguous tiles
    static constexpr int A_ELEMS = (BM * GEMM_LDS_K > GEMM_BK * LDS_A_MN) ? BM * GEMM_LDS_K : GEMM_BK * LDS_A_MN;
    static constexpr int B_ELEMS = (GEMM_BN * GEMM_LDS_K > GEMM_BK * LDS_B_MN) ? GEMM_BN * GEMM_LDS_K : GEMM_BK * LDS_B_MN;
    static constexpr int SMEM_BYTES = STAGES * (A_ELEMS + B_ELEMS) * (int)sizeof(double);
    static constexpr int MIN_CTAS = (BM == 128) ? 1 : 2;
};

// Part q (of GEMM_BK/4) of the cp.async copies of a ROWS x GEMM_BK tile.  The k loop spreads the
// copies of the next stage between its DMMA groups: the LSU is in-order, and a burst of a whole
// stage's LDGSTS at the top of the k-tile delays the fragment LDS queued behind it and starves
// the DMMA pipe (measured on B200: 30.5 -> 33.7 TFLOP/s once the copies are interleaved).
template <int LAYOUT, int ROWS, int THREADS>
__device__ __forceinline__ void gemm_load_tile_part(double* s, const double* g, int64_t ld, int row0,
                                                    int k0, int tid, int q) {
    constexpr int NPARTS = GEMM_BK / 4;
    if (LAYOUT == LAYOUT_K) {
        constexpr int CPR = GEMM_BK / 2;                       // 16-byte chunks per row
        constexpr int PER = (ROWS * CPR) / THREADS / NPARTS;
#pragma unroll
        for (int i = 0; i < PER; i++) {
            int c = tid + (q * PER + i) * THREADS;
            int r = c / CPR, kc = c % CPR;
            cp_async16(s + r * GEMM_LDS_K + 2 * kc, g + (int64_t)(row0 + r) * ld + k0 + 2 * kc);
        }
    } else {
        constexpr int CPR = ROWS / 2;                          // chunks per k-row
        constexpr int PER = (GEMM_BK * CPR) / THREADS / NPARTS;
#pragma unroll
        for (int i = 0; i < PER; i++) {
            int c = tid + (q * PER + i) * THREADS;
            int kr = c / CPR, mc = c % CPR;
            cp_async16(s + kr * (ROWS + 4) + 2 * mc, g + (int64_t)(k0 + kr) * ld + row0 + 2 * mc);
        }
    }
}

template <int AL, int BL, int BM>
__global__ void __launch_bounds__(GemmCfg<BM>::THREADS, GemmCfg<BM>::MIN_CTAS) gemm_dmma_kernel(GemmArgs p) {
    using Cfg = GemmCfg<BM>;
    constexpr int MF = Cfg::MF, NF = Cfg::NF, STAGES = Cfg::STAGES, THREADS = Cfg::THREADS;
    extern __shared__ __align__(16) double gemm_smem[];
    // tile -> (it, jt); heavy tiles first when the k range is triangular
    int it, jt;
    if (p.raster > 0) {
        const int nI = p.M / BM, nJ = p.N / GEMM_BN, G = p.raster;
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
    const int m0 = it * BM, n0 = jt * GEMM_BN;
    if (p.lower_only && n0 > m0 + BM - 1) return;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;

    const double* A = p.A + (int64_t)blockIdx.z * p.strideA;
    const double* B = p.B + (int64_t)blockIdx.z * p.strideB;
    double* C = p.C + (int64_t)blockIdx.z * p.strideC;

    // k range in units of GEMM_BK, from the triangular structure of the operands
    int kt_lo = 0, kt_hi = p.K / GEMM_BK;
    if (p.krange == KR_LE_J) kt_hi = min(kt_hi, (n0 + GEMM_BN) / GEMM_BK);
    else if (p.krange == KR_GE_J) kt_lo = n0 / GEMM_BK;
    else if (p.krange == KR_LE_I) kt_hi = min(kt_hi, (m0 + BM) / GEMM_BK);
    else if (p.krange == KR_GE_I) kt_lo = m0 / GEMM_BK;
    const int nkt = kt_hi - kt_lo;

    double* sA = gemm_smem;
    double* sB = gemm_smem + STAGES * Cfg::A_ELEMS;

    double acc[MF][NF][2];
#pragma unroll
    for (int i = 0; i < MF; i++)
#pragma unroll
        for (int j = 0; j < NF; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nkt) {
#pragma unroll
            for (int q = 0; q < GEMM_BK / 4; q++) {
                gemm_load_tile_part<AL, BM, THREADS>(sA + s * Cfg::A_ELEMS, A, p.lda, m0, (kt_lo + s) * GEMM_BK, tid, q);
                gemm_load_tile_part<BL, GEMM_BN, THREADS>(sB + s * Cfg::B_ELEMS, B, p.ldb, n0, (kt_lo + s) * GEMM_BK, tid, q);
            }
        }
        cp_async_commit();
    }

    // Fragments are double-buffered in registers across k-groups AND across k-tiles: the wait +
    // barrier for the next stage sits before the LAST k-group of the current tile, whose 32 DMMAs
    // then cover the barrier and the first fragment loads of the next tile.
    auto load_frags = [&](const double* a_s, const double* b_s, int kk, double (&af)[MF], double (&bf)[NF]) {
#pragma unroll
        for (int mf = 0; mf < MF; mf++) {
            int m = wm * (MF * 8) + mf * 8 + g, k = kk * 4 + t;
            af[mf] = (AL == LAYOUT_K) ? a_s[m * GEMM_LDS_K + k] : a_s[k * Cfg::LDS_A_MN + m];
        }
#pragma unroll
        for (int nf = 0; nf < NF; nf++) {
            int n = wn * (NF * 8) + nf * 8 + g, k = kk * 4 + t;
            bf[nf] = (BL == LAYOUT_K) ? b_s[n * GEMM_LDS_K + k] : b_s[k * Cfg::LDS_B_MN + n];
        }
    };
    constexpr int NKK = GEMM_BK / 4;
    double af[2][MF], bf[2][NF];
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (nkt > 0) load_frags(sA, sB, 0, af[0], bf[0]);

    for (int kt = 0; kt < nkt; kt++) {
        const int nk = kt + STAGES - 1;
        const bool do_load = nk < nkt;
        double* a_n = sA + (nk % STAGES) * Cfg::A_ELEMS;
        double* b_n = sB + (nk % STAGES) * Cfg::B_ELEMS;
        const int k_n = (kt_lo + nk) * GEMM_BK;
        const double* a_s = sA + (kt % STAGES) * Cfg::A_ELEMS;
        const double* b_s = sB + (kt % STAGES) * Cfg::B_ELEMS;
        const double* a_x = sA + ((kt + 1) % STAGES) * Cfg::A_ELEMS;
        const double* b_x = sB + ((kt + 1) % STAGES) * Cfg::B_ELEMS;
#pragma unroll
        for (int kk = 0; kk < NKK; kk++) {
            const int cur = kk & 1, nxt = cur ^ 1;
            if (kk == NKK - 1) {
                // all copies of this iteration were issued in the earlier k-groups
                cp_async_commit();
                cp_async_wait<STAGES - 2>();
                __syncthreads();
                if (kt + 1 < nkt) load_frags(a_x, b_x, 0, af[nxt], bf[nxt]);
            } else {
                load_frags(a_s, b_s, kk + 1, af[nxt], bf[nxt]);
            }
#pragma unroll
            for (int mf = 0; mf < MF; mf++) {
#pragma unroll
                for (int nf = 0; nf < NF; nf++) dmma884(acc[mf][nf][0], acc[mf][nf][1], af[cur][mf], bf[cur][nf]);
                if (do_load && kk < NKK - 1) {
                    // 4 parts of each operand over the first 3 k-groups: A parts {0,1 | 2 | 3}, same for B
                    if (mf == MF / 4) {
                        gemm_load_tile_part<AL, BM, THREADS>(a_n, A, p.lda, m0, k_n, tid, kk == 0 ? 0 : kk + 1);
                        if (kk == 0) gemm_load_tile_part<AL, BM, THREADS>(a_n, A, p.lda, m0, k_n, tid, 1);
                    }
                    if (mf == (3 * MF) / 4) {
                        gemm_load_tile_part<BL, GEMM_BN, THREADS>(b_n, B, p.ldb, n0, k_n, tid, kk == 0 ? 0 : kk + 1);
                        if (kk == 0) gemm_load_tile_part<BL, GEMM_BN, THREADS>(b_n, B, p.ldb, n0, k_n, tid, 1);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();

    if (p.cmode == C_COLSUMSQ) {
        // column sums of squares over the BM rows of this tile -> partial[it][n0 + col]
        __syncthreads();
        double* red = gemm_smem;   // [WM][128]
#pragma unroll
        for (int nf = 0; nf < NF; nf++) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int mf = 0; mf < MF; mf++) {
                s0 += acc[mf][nf][0] * acc[mf][nf][0];
                s1 += acc[mf][nf][1] * acc[mf][nf][1];
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {     // reduce over g (lane bits 2..4)
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
            if (g == 0) {
                int col = wn * (NF * 8) + nf * 8 + 2 * t;
                red[wm * 128 + col] = s0;
                red[wm * 128 + col + 1] = s1;
            }
        }
        __syncthreads();
        if (tid < 128) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < Cfg::WM; w++) sum += red[w * 128 + tid];
            C[(int64_t)it * p.ldc + n0 + tid] = sum;
        }
        return;
    }

#pragma unroll
    for (int mf = 0; mf < MF; mf++) {
        int row = m0 + wm * (MF * 8) + mf * 8 + g;
#pragma unroll
        for (int nf = 0; nf < NF; nf++) {
            int col = n0 + wn * (NF * 8) + nf * 8 + 2 * t;
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
