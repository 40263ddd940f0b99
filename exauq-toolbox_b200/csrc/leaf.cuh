// Leaf of the recursive factorisation: ONE CTA turns a 128 x 128 SPD diagonal block into
// its Cholesky factor L (lower) and the inverse factor Linv = L^-1, entirely in shared
// memory, with the O(n^3) parts on DMMA.8x8x4.
//
// Replaces (for one diagonal block) what the reference reaches through mogp's
// jit_cholesky -> scipy.linalg.lapack.dpotrf (mogp GaussianProcess.fit via
// exauq/core/emulators.py:447-448); failure (non-positive pivot) is reported through
// `info` so that the host can apply mogp's adaptive-jitter retry policy.
//
// Layout: S[128][132] holds L in the lower triangle; the strictly-upper part is reused to
// hold Linv TRANSPOSED (Linv[r][c], r > c, lives at S[c][r]); the sixteen 8 x 8 diagonal
// blocks of Linv live in XD.  The 136 lower 8 x 8 tiles are owned round-robin by warps 1..7
// and stay in DMMA accumulator registers between steps; warp 0 runs the 8-pivot chain of the
// next diagonal block while they apply the rank-8 update of the current one.
#pragma once
#include "exq_common.cuh"

namespace exq {

#ifdef EXQ_LEAF_PROFILE
__device__ long long leaf_prof[64];
#define LEAF_TICK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) leaf_prof[i] += clock64() - t_prev; t_prev = clock64(); } while (0)
#else
#define LEAF_TICK(i) do { } while (0)
#endif

constexpr int LEAF_N = 128, LEAF_LD = 132, LEAF_NBLK = 16, LEAF_THREADS = 256;
constexpr int LEAF_OWNERS = 7;        // warps 1..7 own the 136 tiles; warp 0 keeps the pivot chain one step ahead
constexpr int LEAF_SLOTS = 20;        // ceil(136 / 7)
constexpr int LEAF_XD_LD = 12;
constexpr int LEAF_SMEM_DOUBLES =
    LEAF_N * LEAF_LD + LEAF_NBLK * 8 * LEAF_XD_LD + LEAF_N + 8 * 8 * LEAF_XD_LD;
constexpr int LEAF_SMEM_BYTES = LEAF_SMEM_DOUBLES * (int)sizeof(double);

// The leaf proper: A and Li point at the 128 x 128 diagonal block (leading dimension ld) of one system,
// *info_sys receives (first bad column + 1 + info_base), leaf_smem is LEAF_SMEM_BYTES of shared memory.
// Called by the stand-alone kernel below and, for CTA 0 of a cluster, by fused_factor_kernel (fused_node.cuh), where
// the block was last written by other CTAs: its global reads bypass L1 (ld.global.cg).
__device__ __forceinline__ void leaf_body(double* __restrict__ A, double* __restrict__ Li, int64_t ld,
                                          int* __restrict__ info_sys, int info_base, double* leaf_smem) {
    double* S = leaf_smem;
    double* XD = S + LEAF_N * LEAF_LD;                 // [16][8][12]
    double* rinv = XD + LEAF_NBLK * 8 * LEAF_XD_LD;    // [128]
    double* scratch = rinv + LEAF_N;                   // [8 warps][8][12]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

#ifdef EXQ_LEAF_PROFILE
    long long t_prev = clock64();
#endif
    // ---- tile ownership -------------------------------------------------------------
    int tb[LEAF_SLOTS];       // (bi << 8) | bj, or -1
    double c[LEAF_SLOTS][2];
#pragma unroll
    for (int s = 0; s < LEAF_SLOTS; s++) {
        // tile q of the column-major enumeration of the 136 lower tiles: column bj starts at off(bj) = 16 bj - bj (bj - 1) / 2
        const int q = (warp - 1) + LEAF_OWNERS * s;
        int bj = (int)((33.0f - sqrtf(1089.0f - 8.0f * (float)min(q, 135))) * 0.5f);
        if (16 * (bj + 1) - ((bj + 1) * bj) / 2 <= q) bj++;
        if (16 * bj - (bj * (bj - 1)) / 2 > q) bj--;
        const int bi = bj + q - (16 * bj - (bj * (bj - 1)) / 2);
        tb[s] = (warp > 0 && q < 136) ? ((bi << 8) | bj) : -1;
    }
    // all 17 tile loads of a thread are issued before the first one is consumed (the index search above used to sit
    // between them and serialised 17 L2 round trips: 12 k of the leaf's 80 k cycles)
    double2 tile_in[LEAF_SLOTS];
#pragma unroll
    for (int s = 0; s < LEAF_SLOTS; s++) {
        tile_in[s] = make_double2(0.0, 0.0);
        if (tb[s] >= 0) {
            const int row = 8 * (tb[s] >> 8) + g, col = 8 * (tb[s] & 255) + 2 * t;
            tile_in[s] = __ldcg(reinterpret_cast<const double2*>(A + (int64_t)row * ld + col));
        }
    }
#pragma unroll
    for (int s = 0; s < LEAF_SLOTS; s++) {
        c[s][0] = c[s][1] = 0.0;
        if (tb[s] >= 0) {
            const int bi = tb[s] >> 8, bj = tb[s] & 255;
            const int row = 8 * bi + g, col = 8 * bj + 2 * t;
            c[s][0] = (col <= row) ? tile_in[s].x : 0.0;
            c[s][1] = (col + 1 <= row) ? tile_in[s].y : 0.0;
            if (bj == 0) {
                S[row * LEAF_LD + col] = c[s][0];
                S[row * LEAF_LD + col + 1] = c[s][1];
            }
        }
    }
    __syncthreads();
    LEAF_TICK(0);

    // ================= phase 1: Cholesky =================================================
    // 8 x 8 diagonal block at j0, warp 0.  Every lane factors the whole block redundantly in
    // registers (36 broadcast loads, no shuffles on the 8-pivot dependency chain).
    auto diag_factor = [&](int j0) {
        double a[8][8];
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int cc = 0; cc <= r; cc++) a[r][cc] = S[(j0 + r) * LEAF_LD + j0 + cc];
        double ri[8];
        bool bad = false;
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {
            const double d = a[cc][cc];
            bad = bad || !(d > 0.0) || !(d < 1.0e300);
            ri[cc] = rsqrt(d);
            a[cc][cc] = d * ri[cc];
#pragma unroll
            for (int r = cc + 1; r < 8; r++) a[r][cc] *= ri[cc];
#pragma unroll
            for (int c2 = cc + 1; c2 < 8; c2++)
#pragma unroll
                for (int r = c2; r < 8; r++) a[r][c2] -= a[r][cc] * a[c2][cc];
        }
        if (lane < 8) {
#pragma unroll
            for (int r = 0; r < 8; r++)
                if (r == lane) {
#pragma unroll
                    for (int cc = 0; cc <= r; cc++) S[(j0 + r) * LEAF_LD + j0 + cc] = a[r][cc];
                    rinv[j0 + r] = ri[r];
                }
        }
        if (lane == 0 && bad) {
            int first = 0;
#pragma unroll
            for (int cc = 7; cc >= 0; cc--)
                if (!(a[cc][cc] > 0.0) || !(a[cc][cc] < 1.0e300) || !(ri[cc] == ri[cc])) first = cc;
            atomicCAS(info_sys, 0, info_base + j0 + first + 1);
        }
    };
    // Step p: (b) panel below diagonal block p, (c1) rank-8 update of block column p + 1 only, then -- one step of
    // look-ahead -- warp 0 factors diagonal block p + 1 WHILE the other warps (and warp 0 afterwards) apply the
    // rank-8 update to the rest of the trailing matrix (c2).  The 8-pivot chain (1400 cycles) used to be a phase of
    // its own with 7 warps idle: 22 k of the leaf's 80 k cycles.
    // Block row i of X = L^-1 (tiles X_ij, j < i), by the tile-owning warps: X_ij = -XD_i * sum_{k=j}^{i-1} L_ik X_kj with
    // X_jj = XD_j.  Row i needs block row i of L, XD_i and the rows above it -- all there once step i's panel phase is
    // done -- so it runs in the shadow of warp 0's pivot chain instead of as a phase of its own after the factorisation
    // (12 k of the leaf's 74 k cycles).  Column j always belongs to the same warp; finished tiles are parked TRANSPOSED in
    // the strictly upper part of S, where that warp reads them back as DMMA B fragments.
    double* my_scratch = scratch + warp * 8 * LEAF_XD_LD;
    auto x_row = [&](int i) {
#pragma unroll 1
        for (int j = warp - 1; j < i; j += LEAF_OWNERS) {
            // Y = sum_k L_ik X_kj, two independent accumulators (even / odd k)
            double y0[2] = {0.0, 0.0}, y1[2] = {0.0, 0.0};
            {
                const double a0 = S[(8 * i + g) * LEAF_LD + 8 * j + t];
                const double a1 = S[(8 * i + g) * LEAF_LD + 8 * j + 4 + t];
                const double b0 = XD[(j * 8 + t) * LEAF_XD_LD + g];
                const double b1 = XD[(j * 8 + 4 + t) * LEAF_XD_LD + g];
                dmma884(y0[0], y0[1], a0, b0);
                dmma884(y1[0], y1[1], a1, b1);
            }
#pragma unroll 2
            for (int k = j + 1; k < i; k++) {
                const double a0 = S[(8 * i + g) * LEAF_LD + 8 * k + t];
                const double a1 = S[(8 * i + g) * LEAF_LD + 8 * k + 4 + t];
                const double b0 = S[(8 * j + g) * LEAF_LD + 8 * k + t];
                const double b1 = S[(8 * j + g) * LEAF_LD + 8 * k + 4 + t];
                dmma884(y0[0], y0[1], a0, b0);
                dmma884(y1[0], y1[1], a1, b1);
            }
            // X_ij = -XD_i * Y : Y goes through scratch to become a B fragment
            my_scratch[g * LEAF_XD_LD + 2 * t] = y0[0] + y1[0];
            my_scratch[g * LEAF_XD_LD + 2 * t + 1] = y0[1] + y1[1];
            __syncwarp();
            const double b0 = my_scratch[t * LEAF_XD_LD + g];
            const double b1 = my_scratch[(4 + t) * LEAF_XD_LD + g];
            const double a0 = XD[(i * 8 + g) * LEAF_XD_LD + t];
            const double a1 = XD[(i * 8 + g) * LEAF_XD_LD + 4 + t];
            double d0 = 0.0, d1 = 0.0;
            dmma884(d0, d1, -a0, b0);
            dmma884(d0, d1, -a1, b1);
            S[(8 * j + 2 * t) * LEAF_LD + 8 * i + g] = d0;
            S[(8 * j + 2 * t + 1) * LEAF_LD + 8 * i + g] = d1;
            __syncwarp();
        }
    };
    if (warp == 0) diag_factor(0);
    __syncthreads();
    LEAF_TICK(1);
    for (int p = 0; p < LEAF_NBLK; p++) {
        const int j0 = 8 * p;
        // (b) rows below the diagonal block: forward substitution with L11; warp 7 inverts L11
        {
            const int r = j0 + 8 + tid;
            if (r < LEAF_N) {
                double a[8];
#pragma unroll
                for (int cc = 0; cc < 8; cc++) a[cc] = S[r * LEAF_LD + j0 + cc];
#pragma unroll
                for (int cc = 0; cc < 8; cc++) {
                    double s = a[cc];
#pragma unroll
                    for (int k = 0; k < cc; k++) s -= a[k] * S[(j0 + cc) * LEAF_LD + j0 + k];
                    a[cc] = s * rinv[j0 + cc];
                }
#pragma unroll
                for (int cc = 0; cc < 8; cc++) S[r * LEAF_LD + j0 + cc] = a[cc];
            }
            if (warp == 7 && lane < 8) {
                // column `lane` of inv(L11)
                const int cc = lane;
                double x[8];
#pragma unroll
                for (int rr = 0; rr < 8; rr++) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < rr; k++)
                        if (k >= cc) s += S[(j0 + rr) * LEAF_LD + j0 + k] * x[k];
                    x[rr] = (rr < cc) ? 0.0 : ((rr == cc) ? rinv[j0 + rr] : -s * rinv[j0 + rr]);
                }
#pragma unroll
                for (int rr = 0; rr < 8; rr++) XD[(p * 8 + rr) * LEAF_XD_LD + cc] = x[rr];
            }
        }
        __syncthreads();
        LEAF_TICK(2);
        // (c) rank-8 update of every trailing tile, in registers.  Tiles are enumerated column-major,
        // so the tiles with bj > p are a contiguous SUFFIX of this warp's slots: enter the unrolled
        // slot sequence at the first active one (no per-slot branch, loads and DMMAs of successive
        // slots overlap).  Block column p + 1 (tiles q0 <= q < q1) goes first and is published to S.
        const int q0 = 16 * (p + 1) - ((p + 1) * p) / 2;          // first tile of block column p+1
        const int q1 = 16 * (p + 2) - ((p + 2) * (p + 1)) / 2;    // first tile of block column p+2
#define LEAF_SLOT_UPDATE(SL, FIRST_COLUMN)                                                \
    if (tb[SL] >= 0 && (((warp - 1) + LEAF_OWNERS * SL < q1) == (FIRST_COLUMN))) {       \
        const int bi = tb[SL] >> 8, bj = tb[SL] & 255;                                    \
        double a0 = S[(8 * bi + g) * LEAF_LD + j0 + t];                                   \
        double a1 = S[(8 * bi + g) * LEAF_LD + j0 + 4 + t];                               \
        double b0 = S[(8 * bj + g) * LEAF_LD + j0 + t];                                   \
        double b1 = S[(8 * bj + g) * LEAF_LD + j0 + 4 + t];                               \
        dmma884(c[SL][0], c[SL][1], -a0, b0);                                             \
        dmma884(c[SL][0], c[SL][1], -a1, b1);                                             \
        if (FIRST_COLUMN) {                                                               \
            int row = 8 * bi + g, col = 8 * bj + 2 * t;                                   \
            S[row * LEAF_LD + col] = c[SL][0];                                            \
            S[row * LEAF_LD + col + 1] = c[SL][1];                                        \
        }                                                                                 \
    }
#define LEAF_ALL_SLOTS(FROM, FIRST_COLUMN)                                                \
    switch (FROM) {                                                                       \
        case 0: LEAF_SLOT_UPDATE(0, FIRST_COLUMN)                                         \
        case 1: LEAF_SLOT_UPDATE(1, FIRST_COLUMN)                                         \
        case 2: LEAF_SLOT_UPDATE(2, FIRST_COLUMN)                                         \
        case 3: LEAF_SLOT_UPDATE(3, FIRST_COLUMN)                                         \
        case 4: LEAF_SLOT_UPDATE(4, FIRST_COLUMN)                                         \
        case 5: LEAF_SLOT_UPDATE(5, FIRST_COLUMN)                                         \
        case 6: LEAF_SLOT_UPDATE(6, FIRST_COLUMN)                                         \
        case 7: LEAF_SLOT_UPDATE(7, FIRST_COLUMN)                                         \
        case 8: LEAF_SLOT_UPDATE(8, FIRST_COLUMN)                                         \
        case 9: LEAF_SLOT_UPDATE(9, FIRST_COLUMN)                                         \
        case 10: LEAF_SLOT_UPDATE(10, FIRST_COLUMN)                                       \
        case 11: LEAF_SLOT_UPDATE(11, FIRST_COLUMN)                                       \
        case 12: LEAF_SLOT_UPDATE(12, FIRST_COLUMN)                                       \
        case 13: LEAF_SLOT_UPDATE(13, FIRST_COLUMN)                                       \
        case 14: LEAF_SLOT_UPDATE(14, FIRST_COLUMN)                                       \
        case 15: LEAF_SLOT_UPDATE(15, FIRST_COLUMN)                                       \
        case 16: LEAF_SLOT_UPDATE(16, FIRST_COLUMN)                                       \
        case 17: LEAF_SLOT_UPDATE(17, FIRST_COLUMN)                                       \
        case 18: LEAF_SLOT_UPDATE(18, FIRST_COLUMN)                                       \
        case 19: LEAF_SLOT_UPDATE(19, FIRST_COLUMN)                                       \
        default: break;                                                                   \
    }
        // first slot of this warp with q >= q0 / q >= q1 (q = warp - 1 + 7 s)
        const int smin = (q0 - (warp - 1) + LEAF_OWNERS - 1) / LEAF_OWNERS;
        const int smin2 = (q1 - (warp - 1) + LEAF_OWNERS - 1) / LEAF_OWNERS;
        // (c1) block column p + 1: at most three slots per warp; later slots fail the (q < q1) test at once
        if (warp > 0) { LEAF_ALL_SLOTS(smin < 0 ? 0 : smin, true) }
        __syncthreads();
        // (a, one step ahead) by warp 0, which owns no tiles, beside (c2) by warps 1..7
        if (warp == 0) {
            if (p + 1 < LEAF_NBLK) diag_factor(j0 + 8);
        } else {
            LEAF_ALL_SLOTS(smin2 < 0 ? 0 : smin2, false)
            if (p > 0) x_row(p);
        }
#undef LEAF_ALL_SLOTS
#undef LEAF_SLOT_UPDATE
        __syncthreads();
        LEAF_TICK(3);
    }

    LEAF_TICK(5);
    __syncthreads();

    // ---- write back L (upper zeroed) and Linv (full tile, zeros above the diagonal), 16 bytes per store ----
    for (int idx = tid; idx < LEAF_N * (LEAF_N / 2); idx += LEAF_THREADS) {
        const int r = idx >> 6, cc = (idx & 63) << 1;
        double2 l;
        l.x = (cc <= r) ? S[r * LEAF_LD + cc] : 0.0;
        l.y = (cc + 1 <= r) ? S[r * LEAF_LD + cc + 1] : 0.0;
        *reinterpret_cast<double2*>(A + (int64_t)r * ld + cc) = l;
    }
    // Linv tile by tile (a warp per 8 x 8 tile, lane (g, t) holds row g, columns 2t, 2t+1): the parked tiles are read
    // along their shared-memory rows (a row-major sweep of the output read them with a stride of 2 LD doubles: 16-way
    // bank conflicts, half of the write-back's 8.9 k cycles)
    for (int tile = warp; tile < LEAF_NBLK * LEAF_NBLK; tile += LEAF_THREADS / 32) {
        const int br = tile >> 4, bc = tile & 15;
        double2 x;
        if (bc < br) {
            x.x = S[(8 * bc + 2 * t) * LEAF_LD + 8 * br + g];
            x.y = S[(8 * bc + 2 * t + 1) * LEAF_LD + 8 * br + g];
        } else if (bc == br) {
            x.x = XD[(br * 8 + g) * LEAF_XD_LD + 2 * t];
            x.y = XD[(br * 8 + g) * LEAF_XD_LD + 2 * t + 1];
        } else {
            x.x = x.y = 0.0;
        }
        *reinterpret_cast<double2*>(Li + (int64_t)(8 * br + g) * ld + 8 * bc + 2 * t) = x;
    }
    LEAF_TICK(6);
    __syncthreads();   // shared memory is free again for the caller
}

// A, Linv: pointers to the (row-major, leading dimension ld) full matrices of batch 0;
// the leaf works on rows/cols [r0, r0+128).  info[batch] receives first bad column + 1.
__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_trtri_kernel(double* __restrict__ Aall, double* __restrict__ Linvall, int64_t ld,
                        int64_t strideA, int64_t strideL, int r0, int* __restrict__ info) {
    extern __shared__ __align__(16) double leaf_smem_dyn[];
    leaf_body(Aall + (int64_t)blockIdx.x * strideA + (int64_t)r0 * ld + r0,
              Linvall + (int64_t)blockIdx.x * strideL + (int64_t)r0 * ld + r0, ld, info + blockIdx.x, r0, leaf_smem_dyn);
}

}  // namespace exq
