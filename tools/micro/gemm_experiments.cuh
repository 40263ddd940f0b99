// Experimental GEMM variants kept for the record (not used by the product): a warp-specialised
// mbarrier pipeline with a cp.async producer warp, and TMA tensor-map fed variants.  Measured on
// B200 (tools/micro/gemm_bench.cu, variance shape Np=4096, Mc=32768): classic burst-issue ring
// 30.5, warp-specialised 30.5, TMA {4,128} boxes 32.3, TMA 128B-swizzle box 31.6, classic with
// cp.async interleaved between DMMA groups 33.7 TFLOP/s (cuBLAS 36.4, DMMA issue peak 37.0).
#pragma once
#include "../../exauq-toolbox_b200/csrc/gemm_dmma.cuh"
namespace exq {
// constants of the 128 x 128 configuration these variants were written against
constexpr int GEMM_BM = 128, GEMM_STAGES = 4, GEMM_WM = 2, GEMM_WN = 4, GEMM_THREADS = 256;
constexpr int GEMM_MF = GEMM_BM / GEMM_WM / 8, GEMM_NF = GEMM_BN / GEMM_WN / 8;
constexpr int GEMM_LDS_MN = GEMM_BM + 4;
constexpr int GEMM_TILE_ELEMS = GEMM_BM * GEMM_LDS_K;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * 2 * GEMM_TILE_ELEMS * (int)sizeof(double);
// ---------------------------------------------------------------------------------------
// Warp-specialised variant: 8 consumer warps (LDS + DMMA only) and 1 producer warp that
// issues every cp.async of the CTA.  Stages are handed over through mbarriers
// (full[s]: 32 pre-counted arrivals released by cp.async.mbarrier.arrive when the
// producer lanes' copies have landed; empty[s]: one arrival per consumer warp), so there
// is no CTA-wide barrier in the k loop and consumer warps drift freely.  Measured on B200
// (tools/micro/gemm_bench.cu): the __syncthreads + per-warp cp.async issue of the classic
// ring cost 14 % of the DMMA rate (30.5 vs 35.4 TFLOP/s with loads and barriers removed).
// ---------------------------------------------------------------------------------------
constexpr int GEMM_WS_THREADS = GEMM_THREADS + 32;
constexpr int GEMM_WS_SMEM_BYTES = GEMM_SMEM_BYTES + 2 * GEMM_STAGES * (int)sizeof(uint64_t);

template <int LAYOUT>
__device__ __forceinline__ void gemm_load_tile_warp(double* s, const double* g, int64_t ld, int row0,
                                                    int k0, int lane) {
    if (LAYOUT == LAYOUT_K) {
        constexpr int CPR = GEMM_BK / 2;
#pragma unroll 8
        for (int i = 0; i < (GEMM_BM * CPR) / 32; i++) {
            int c = lane + i * 32;
            int r = c / CPR, kc = c % CPR;
            cp_async16(s + r * GEMM_LDS_K + 2 * kc, g + (int64_t)(row0 + r) * ld + k0 + 2 * kc);
        }
    } else {
#pragma unroll 8
        for (int i = 0; i < (GEMM_BK * 64) / 32; i++) {
            int c = lane + i * 32;
            int kr = c >> 6, mc = c & 63;
            cp_async16(s + kr * GEMM_LDS_MN + 2 * mc, g + (int64_t)(k0 + kr) * ld + row0 + 2 * mc);
        }
    }
}

// TMA (cp.async.bulk) flavour of the producer: one bulk copy per contiguous row segment of the
// tile, all completing on the stage's mbarrier through complete_tx::bytes.
template <int LAYOUT>
__device__ __forceinline__ void gemm_load_tile_tma(double* s, const double* g, int64_t ld, int row0,
                                                   int k0, int lane, uint64_t* bar) {
    if (LAYOUT == LAYOUT_K) {
#pragma unroll
        for (int i = 0; i < GEMM_BM / 32; i++) {
            int r = lane + i * 32;
            tma_bulk_g2s(s + r * GEMM_LDS_K, g + (int64_t)(row0 + r) * ld + k0, GEMM_BK * sizeof(double), bar);
        }
    } else {
        if (lane < GEMM_BK)
            tma_bulk_g2s(s + lane * GEMM_LDS_MN, g + (int64_t)(k0 + lane) * ld + row0, GEMM_BM * sizeof(double), bar);
    }
}

template <int AL, int BL>
__global__ void __launch_bounds__(GEMM_WS_THREADS, 1) gemm_dmma_ws_kernel(GemmArgs p) {
    extern __shared__ __align__(16) double gemm_smem[];
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

    const double* A = p.A + (int64_t)blockIdx.z * p.strideA;
    const double* B = p.B + (int64_t)blockIdx.z * p.strideB;
    double* C = p.C + (int64_t)blockIdx.z * p.strideC;

    int kt_lo = 0, kt_hi = p.K / GEMM_BK;
    const int tiles_per_blk = GEMM_BM / GEMM_BK;
    if (p.krange == KR_LE_J) kt_hi = min(kt_hi, (jt + 1) * tiles_per_blk);
    else if (p.krange == KR_GE_J) kt_lo = jt * tiles_per_blk;
    else if (p.krange == KR_LE_I) kt_hi = min(kt_hi, (it + 1) * tiles_per_blk);
    else if (p.krange == KR_GE_I) kt_lo = it * tiles_per_blk;
    const int nkt = kt_hi - kt_lo;

    double* sA = gemm_smem;
    double* sB = gemm_smem + GEMM_STAGES * GEMM_TILE_ELEMS;
    uint64_t* full = reinterpret_cast<uint64_t*>(gemm_smem + 2 * GEMM_STAGES * GEMM_TILE_ELEMS);
    uint64_t* empty = full + GEMM_STAGES;
    const int m0 = it * GEMM_BM, n0 = jt * GEMM_BN;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GEMM_STAGES; s++) {
#ifdef EXQ_GEMM_TMA_PRODUCER
            mbar_init(&full[s], 1);
#else
            mbar_init(&full[s], 32);
#endif
            mbar_init(&empty[s], GEMM_WM * GEMM_WN);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == GEMM_WM * GEMM_WN) {
        // ===== producer warp =====
        for (int kt = 0; kt < nkt; kt++) {
            const int s = kt % GEMM_STAGES;
            if (kt >= GEMM_STAGES) mbar_wait(&empty[s], ((kt / GEMM_STAGES) - 1) & 1);
#ifdef EXQ_GEMM_TMA_PRODUCER
            if (lane == 0) mbar_expect_tx(&full[s], 2u * GEMM_BM * GEMM_BK * sizeof(double));
            __syncwarp();
            gemm_load_tile_tma<AL>(sA + s * GEMM_TILE_ELEMS, A, p.lda, m0, (kt_lo + kt) * GEMM_BK, lane, &full[s]);
            gemm_load_tile_tma<BL>(sB + s * GEMM_TILE_ELEMS, B, p.ldb, n0, (kt_lo + kt) * GEMM_BK, lane, &full[s]);
#else
            gemm_load_tile_warp<AL>(sA + s * GEMM_TILE_ELEMS, A, p.lda, m0, (kt_lo + kt) * GEMM_BK, lane);
#if defined(EXQ_GEMM_EXPERIMENT_HALF)
            // only A is loaded
#elif defined(EXQ_GEMM_EXPERIMENT_SAMEADDR)
            gemm_load_tile_warp<BL>(sB + s * GEMM_TILE_ELEMS, B, p.ldb, 0, 0, lane);
#else
            gemm_load_tile_warp<BL>(sB + s * GEMM_TILE_ELEMS, B, p.ldb, n0, (kt_lo + kt) * GEMM_BK, lane);
#endif
            cp_async_mbar_arrive_noinc(&full[s]);
#endif
        }
        cp_async_wait<0>();
        return;
    }

    // ===== consumer warps =====
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / GEMM_WN, wn = warp % GEMM_WN;
    double acc[GEMM_MF][GEMM_NF][2];
#pragma unroll
    for (int i = 0; i < GEMM_MF; i++)
#pragma unroll
        for (int j = 0; j < GEMM_NF; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < nkt; kt++) {
        const int s = kt % GEMM_STAGES;
#ifndef EXQ_GEMM_EXPERIMENT_NOWAIT
        mbar_wait(&full[s], (kt / GEMM_STAGES) & 1);
#endif
        const double* a_s = sA + s * GEMM_TILE_ELEMS;
        const double* b_s = sB + s * GEMM_TILE_ELEMS;
#pragma unroll
        for (int kk = 0; kk < GEMM_BK / 4; kk++) {
            double af[GEMM_MF], bf[GEMM_NF];
#pragma unroll
            for (int mf = 0; mf < GEMM_MF; mf++) {
                int m = wm * (GEMM_MF * 8) + mf * 8 + g, k = kk * 4 + t;
                af[mf] = (AL == LAYOUT_K) ? a_s[m * GEMM_LDS_K + k] : a_s[k * GEMM_LDS_MN + m];
            }
#pragma unroll
            for (int nf = 0; nf < GEMM_NF; nf++) {
                int n = wn * (GEMM_NF * 8) + nf * 8 + g, k = kk * 4 + t;
                bf[nf] = (BL == LAYOUT_K) ? b_s[n * GEMM_LDS_K + k] : b_s[k * GEMM_LDS_MN + n];
            }
#pragma unroll
            for (int mf = 0; mf < GEMM_MF; mf++)
#pragma unroll
                for (int nf = 0; nf < GEMM_NF; nf++) dmma884(acc[mf][nf][0], acc[mf][nf][1], af[mf], bf[nf]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    if (p.cmode == C_COLSUMSQ) {
        named_bar_sync(1, GEMM_THREADS);           // all consumers are done with the stage buffers
        double* red = gemm_smem;                   // [GEMM_WM][128]
#pragma unroll
        for (int nf = 0; nf < GEMM_NF; nf++) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int mf = 0; mf < GEMM_MF; mf++) {
                s0 += acc[mf][nf][0] * acc[mf][nf][0];
                s1 += acc[mf][nf][1] * acc[mf][nf][1];
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
            if (g == 0) {
                int col = wn * (GEMM_NF * 8) + nf * 8 + 2 * t;
                red[wm * 128 + col] = s0;
                red[wm * 128 + col + 1] = s1;
            }
        }
        named_bar_sync(1, GEMM_THREADS);
        if (tid < 128) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < GEMM_WM; w++) sum += red[w * 128 + tid];
            C[(int64_t)it * p.ldc + n0 + tid] = sum;
        }
        return;
    }

#pragma unroll
    for (int mf = 0; mf < GEMM_MF; mf++) {
        int row = m0 + wm * (GEMM_MF * 8) + mf * 8 + g;
#pragma unroll
        for (int nf = 0; nf < GEMM_NF; nf++) {
            int col = n0 + wn * (GEMM_NF * 8) + nf * 8 + 2 * t;
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

// ---------------------------------------------------------------------------------------
// TMA-fed variant (both operands k-contiguous).  Each stage is filled by 2 * BK/4 tiled TMA
// loads (cp.async.bulk.tensor -> UTMALDG) of boxes {4 k, 128 rows}: a box lands densely as
// [row][4], so a stage is laid out [k/4][row][4] and the DMMA fragment address
// (k/4)*512 + row*4 + k%4 is bank-conflict free with no padding and no swizzle.
// One elected producer lane issues the loads; 8 consumer warps do LDS + DMMA only.
// ---------------------------------------------------------------------------------------
#ifndef EXQ_GEMM_TMA_SWIZZLE
#define EXQ_GEMM_TMA_SWIZZLE 1   // 1: one {16 k, 128 rows} box per operand with 128-byte swizzle; 0: {4,128} boxes
#endif
constexpr int GEMM_TMA_STAGE_ELEMS = GEMM_BM * GEMM_BK;          // per operand, dense
constexpr int GEMM_TMA_SMEM_BYTES =
    GEMM_STAGES * 2 * GEMM_TMA_STAGE_ELEMS * (int)sizeof(double) + 2 * GEMM_STAGES * (int)sizeof(uint64_t) + 1024;

__global__ void __launch_bounds__(GEMM_WS_THREADS, 1)
gemm_dmma_tma_kernel(GemmArgs p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
    extern __shared__ __align__(128) double gemm_smem[];
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
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* C = p.C + (int64_t)blockIdx.z * p.strideC;
    int kt_lo = 0, kt_hi = p.K / GEMM_BK;
    const int tiles_per_blk = GEMM_BM / GEMM_BK;
    if (p.krange == KR_LE_J) kt_hi = min(kt_hi, (jt + 1) * tiles_per_blk);
    else if (p.krange == KR_GE_J) kt_lo = jt * tiles_per_blk;
    else if (p.krange == KR_LE_I) kt_hi = min(kt_hi, (it + 1) * tiles_per_blk);
    else if (p.krange == KR_GE_I) kt_lo = it * tiles_per_blk;
    const int nkt = kt_hi - kt_lo;

    // 128-byte aligned stage buffers
    double* base = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(gemm_smem) + 1023) & ~uintptr_t(1023));
    double* sA = base;
    double* sB = base + GEMM_STAGES * GEMM_TMA_STAGE_ELEMS;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + 2 * GEMM_STAGES * GEMM_TMA_STAGE_ELEMS);
    uint64_t* empty = full + GEMM_STAGES;
    const int m0 = it * GEMM_BM, n0 = jt * GEMM_BN;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GEMM_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GEMM_WM * GEMM_WN);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == GEMM_WM * GEMM_WN) {
        if (lane == 0) {
            for (int kt = 0; kt < nkt; kt++) {
                const int s = kt % GEMM_STAGES;
                if (kt >= GEMM_STAGES) mbar_wait(&empty[s], ((kt / GEMM_STAGES) - 1) & 1);
                mbar_expect_tx(&full[s], 2u * GEMM_TMA_STAGE_ELEMS * sizeof(double));
                const int k0 = (kt_lo + kt) * GEMM_BK;
#if EXQ_GEMM_TMA_SWIZZLE
                tma_load_3d(sA + s * GEMM_TMA_STAGE_ELEMS, &mapA, k0, m0, blockIdx.z, &full[s]);
                tma_load_3d(sB + s * GEMM_TMA_STAGE_ELEMS, &mapB, k0, n0, blockIdx.z, &full[s]);
#else
#pragma unroll
                for (int q = 0; q < GEMM_BK / 4; q++) {
                    tma_load_3d(sA + s * GEMM_TMA_STAGE_ELEMS + q * (GEMM_BM * 4), &mapA, k0 + 4 * q, m0, blockIdx.z, &full[s]);
                    tma_load_3d(sB + s * GEMM_TMA_STAGE_ELEMS + q * (GEMM_BN * 4), &mapB, k0 + 4 * q, n0, blockIdx.z, &full[s]);
                }
#endif
            }
        }
        return;
    }

    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / GEMM_WN, wn = warp % GEMM_WN;
    double acc[GEMM_MF][GEMM_NF][2];
#pragma unroll
    for (int i = 0; i < GEMM_MF; i++)
#pragma unroll
        for (int j = 0; j < GEMM_NF; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < nkt; kt++) {
        const int s = kt % GEMM_STAGES;
        mbar_wait(&full[s], (kt / GEMM_STAGES) & 1);
#if EXQ_GEMM_TMA_SWIZZLE
        // row r of the tile occupies 128 B; its 16-byte chunk c sits at chunk c ^ (r & 7).  Fragment row
        // g is mapped to tile row pg = 2*(g%4) + g/4 inside each group of 8 so that a half-warp touches
        // rows {0,2,4,6} or {1,3,5,7}: the 16 lanes then hit 16 distinct 8-byte bank pairs.
        const int pg = ((g & 3) << 1) | (g >> 2);
        const double* a_s = sA + s * GEMM_TMA_STAGE_ELEMS + (wm * (GEMM_MF * 8) + pg) * 16 + (t & 1);
        const double* b_s = sB + s * GEMM_TMA_STAGE_ELEMS + (wn * (GEMM_NF * 8) + pg) * 16 + (t & 1);
#else
        const double* a_s = sA + s * GEMM_TMA_STAGE_ELEMS + (wm * (GEMM_MF * 8) + g) * 4 + t;
        const double* b_s = sB + s * GEMM_TMA_STAGE_ELEMS + (wn * (GEMM_NF * 8) + g) * 4 + t;
#endif
#pragma unroll
        for (int kk = 0; kk < GEMM_BK / 4; kk++) {
            double af[GEMM_MF], bf[GEMM_NF];
#if EXQ_GEMM_TMA_SWIZZLE
            const int ch = (((kk << 1) | (t >> 1)) ^ pg) << 1;
#pragma unroll
            for (int mf = 0; mf < GEMM_MF; mf++) af[mf] = a_s[mf * 128 + ch];
#pragma unroll
            for (int nf = 0; nf < GEMM_NF; nf++) bf[nf] = b_s[nf * 128 + ch];
#else
#pragma unroll
            for (int mf = 0; mf < GEMM_MF; mf++) af[mf] = a_s[kk * (GEMM_BM * 4) + mf * 32];
#pragma unroll
            for (int nf = 0; nf < GEMM_NF; nf++) bf[nf] = b_s[kk * (GEMM_BN * 4) + nf * 32];
#endif
#pragma unroll
            for (int mf = 0; mf < GEMM_MF; mf++)
#pragma unroll
                for (int nf = 0; nf < GEMM_NF; nf++) dmma884(acc[mf][nf][0], acc[mf][nf][1], af[mf], bf[nf]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    if (p.cmode == C_COLSUMSQ) {
        named_bar_sync(1, GEMM_THREADS);
        double* red = base;
#pragma unroll
        for (int nf = 0; nf < GEMM_NF; nf++) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int mf = 0; mf < GEMM_MF; mf++) {
                s0 += acc[mf][nf][0] * acc[mf][nf][0];
                s1 += acc[mf][nf][1] * acc[mf][nf][1];
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
            if (g == 0) {
                int col = wn * (GEMM_NF * 8) + nf * 8 + 2 * t;
                red[wm * 128 + col] = s0;
                red[wm * 128 + col + 1] = s1;
            }
        }
        named_bar_sync(1, GEMM_THREADS);
        if (tid < 128) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < GEMM_WM; w++) sum += red[w * 128 + tid];
            C[(int64_t)it * p.ldc + n0 + tid] = sum;
        }
        return;
    }
#pragma unroll
    for (int mf = 0; mf < GEMM_MF; mf++) {
        int row = m0 + wm * (GEMM_MF * 8) + mf * 8 + g;
#pragma unroll
        for (int nf = 0; nf < GEMM_NF; nf++) {
            int col = n0 + wn * (GEMM_NF * 8) + nf * 8 + 2 * t;
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
