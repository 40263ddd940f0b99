// One launch for a whole sub-tree of the potrf + trtri recursion (diagonal blocks of up to 512 or 1024 rows):
// a thread-block CLUSTER per system walks a static list of steps -- 128 x 128 leaves (leaf.cuh) and the four
// tile GEMMs of every internal node (DMMA.8x8x4) -- with cluster barriers where the separate kernels had launch
// boundaries.  The matrices stay in global memory (L2-resident: a 512 x 512 block of A and of Linv is 4 MB per
// system); what the cluster removes is the latency of ~30 dependent launches per 512-block (33 us per GEMM launch
// and 50 us per leaf with 15 of 148 SMs busy: 150 ms of a 676 ms design batch in round 1, and the whole
// strong-scaling floor), and it lets step (3) of a node run beside the next leaf.
//
// Replaces, for a diagonal block, what the reference reaches through mogp's jit_cholesky -> LAPACK dpotrf
// (exauq/core/emulators.py:447-448) plus the triangular inverse the engine keeps instead of cho_solve.
//
//   node(r0, n):  node(A11);  (1) L21 = A21 Linv11^T;  (2) A22 -= L21 L21^T;  (3) T = L21 Linv11;
//                 node(A22);  (4) Linv21 = -Linv22 T            -- same schedule as factor_rec in exq_api.cu
//
// Data written by one CTA and read by another inside the kernel goes through L2: operand tiles are fetched with
// cp.async.cg, scalar reads use ld.global.cg, every step ends with __threadfence + barrier.cluster.
#pragma once
#include <cooperative_groups.h>

#include "gemm_dmma.cuh"
#include "leaf.cuh"

namespace exq {

#ifdef EXQ_FUSED_PROFILE
// globaltimer stamps of system 0: [cluster rank][step][0: own work done, 1: past the barrier]  (tools/micro/fused_bench.cu)
__device__ unsigned long long fused_stamps[16][48][2];
__device__ __forceinline__ unsigned long long fused_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define FUSED_STAMP(I, W) do { if (threadIdx.x == 0 && blockIdx.y == 0) fused_stamps[rank][I][W] = fused_now(); } while (0)
#else
#define FUSED_STAMP(I, W) do { } while (0)
#endif

enum : int { FOP_LEAF = 0, FOP_GEMM = 1 };
enum : int { FT_32x32 = 0, FT_32x64 = 1, FT_64x128 = 2, FT_128x128 = 3 };

struct FusedOp {
    int type;            // FOP_*
    int r0;              // leaf: first row / column of the diagonal block
    // GEMM: C (op)= A B^T-like product, operands addressed inside the system's A (mat 0) or Linv (mat 1) matrix
    int a_mat, b_mat, c_mat;
    int a_off_r, a_off_c, b_off_r, b_off_c, c_off_r, c_off_c;
    int M, N, K;
    int b_layout;        // LAYOUT_K / LAYOUT_MN (A is always LAYOUT_K in the recursion)
    int krange, lower_only, cmode;
    int tile;            // FT_*
    int sync_after;      // cluster barrier after this step (0: the next step is independent of it)
    int skip_rank0;      // CTA 0 takes no tile: it runs the leaf that follows, concurrently
};

constexpr int FUSED_MAX_OPS = 40;          // a 1024-block: 8 leaves + 7 nodes x 4
constexpr int FUSED_THREADS = 256;
// cp.async stages per tile shape: as many as fit the budget, at most 16 (the whole k extent of a 256-wide node).  With
// three stages every 16-wide k tile of a step waited for its own L2 round trip: a 128 x 128 x 128 step took 6-7 us of
// which 2 us are DMMA time (tools/micro/fused_bench.cu).
constexpr int FUSED_GEMM_BUDGET = 200 * 1024;

struct FusedArgs {
    double* A;
    double* Linv;
    int64_t ld, stride;                    // leading dimension and system stride of both matrices
    int* info;                             // [B]
    int n_ops;
    FusedOp ops[FUSED_MAX_OPS];
};

constexpr int fused_tile_bm(int t) { return t == FT_128x128 ? 128 : t == FT_64x128 ? 64 : 32; }
constexpr int fused_tile_bn(int t) { return t == FT_128x128 || t == FT_64x128 ? 128 : t == FT_32x64 ? 64 : 32; }
__host__ __device__ constexpr int fused_stage_bytes(int bm, int bn) {
    // A tile bm x 16 (k-contiguous, padded) + B tile: bn x 16 padded or 16 x (bn + 4): the larger of the two layouts
    return (bm * GEMM_LDS_K + (bn * GEMM_LDS_K > GEMM_BK * (bn + 4) ? bn * GEMM_LDS_K : GEMM_BK * (bn + 4))) * (int)sizeof(double);
}
__host__ __device__ constexpr int fused_stages(int bm, int bn) {
    return FUSED_GEMM_BUDGET / fused_stage_bytes(bm, bn) > 16 ? 16 : FUSED_GEMM_BUDGET / fused_stage_bytes(bm, bn);
}
constexpr int FUSED_SMEM_BYTES = (LEAF_SMEM_BYTES > FUSED_GEMM_BUDGET ? LEAF_SMEM_BYTES : FUSED_GEMM_BUDGET);

// One BM x BN output tile with all 256 threads (8 warps 2 x 4), k range [kt_lo, kt_hi) in units of 16.
template <int BL, int BM, int BN>
__device__ __forceinline__ void fused_gemm_tile(const double* __restrict__ A, const double* __restrict__ B,
                                                double* __restrict__ C, int64_t ld, int m0, int n0, int kt_lo, int kt_hi,
                                                int cmode, double* smem) {
    constexpr int WN = 4, MF = BM / 2 / 8, NF = BN / WN / 8;
    constexpr int FUSED_STAGES = fused_stages(BM, BN);
    static_assert(FUSED_STAGES >= 3, "tile too large for the shared-memory budget");
    constexpr int A_ELEMS = BM * GEMM_LDS_K;
    constexpr int B_LD_MN = BN + 4;
    constexpr int B_ELEMS = (BL == LAYOUT_K) ? BN * GEMM_LDS_K : GEMM_BK * B_LD_MN;
    static_assert(MF >= 1 && NF >= 1, "tile too small for 8 warps");
    double* sA = smem;
    double* sB = smem + FUSED_STAGES * A_ELEMS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / WN, wn = warp % WN;
    const int nkt = kt_hi - kt_lo;

    double acc[MF][NF][2];
#pragma unroll
    for (int i = 0; i < MF; i++)
#pragma unroll
        for (int j = 0; j < NF; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int st, int kt) {
        const int k0 = kt * GEMM_BK;
        double* a_s = sA + st * A_ELEMS;
        double* b_s = sB + st * B_ELEMS;
        for (int c = tid; c < BM * (GEMM_BK / 2); c += FUSED_THREADS) {
            const int r = c / (GEMM_BK / 2), kc = c % (GEMM_BK / 2);
            cp_async16(a_s + r * GEMM_LDS_K + 2 * kc, A + (int64_t)(m0 + r) * ld + k0 + 2 * kc);
        }
        if (BL == LAYOUT_K) {
            for (int c = tid; c < BN * (GEMM_BK / 2); c += FUSED_THREADS) {
                const int r = c / (GEMM_BK / 2), kc = c % (GEMM_BK / 2);
                cp_async16(b_s + r * GEMM_LDS_K + 2 * kc, B + (int64_t)(n0 + r) * ld + k0 + 2 * kc);
            }
        } else {
            for (int c = tid; c < GEMM_BK * (BN / 2); c += FUSED_THREADS) {
                const int kr = c / (BN / 2), nc = c % (BN / 2);
                cp_async16(b_s + kr * B_LD_MN + 2 * nc, B + (int64_t)(k0 + kr) * ld + n0 + 2 * nc);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < FUSED_STAGES - 1; s++) {
        if (s < nkt) load_stage(s, kt_lo + s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nkt; kt++) {
        cp_async_wait<FUSED_STAGES - 2>();
        __syncthreads();                              // stage kt has landed; everyone is done with stage kt - 1
        const int nk = kt + FUSED_STAGES - 1;
        if (nk < nkt) load_stage(nk % FUSED_STAGES, kt_lo + nk);
        cp_async_commit();
        const double* a_s = sA + (kt % FUSED_STAGES) * A_ELEMS;
        const double* b_s = sB + (kt % FUSED_STAGES) * B_ELEMS;
#pragma unroll
        for (int kk = 0; kk < GEMM_BK / 4; kk++) {
            double af[MF], bf[NF];
#pragma unroll
            for (int mf = 0; mf < MF; mf++) af[mf] = a_s[(wm * (MF * 8) + mf * 8 + g) * GEMM_LDS_K + kk * 4 + t];
#pragma unroll
            for (int nf = 0; nf < NF; nf++) {
                const int n = wn * (NF * 8) + nf * 8 + g, k = kk * 4 + t;
                bf[nf] = (BL == LAYOUT_K) ? b_s[n * GEMM_LDS_K + k] : b_s[k * B_LD_MN + n];
            }
#pragma unroll
            for (int mf = 0; mf < MF; mf++)
#pragma unroll
                for (int nf = 0; nf < NF; nf++) dmma884(acc[mf][nf][0], acc[mf][nf][1], af[mf], bf[nf]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mf = 0; mf < MF; mf++) {
        const int row = m0 + wm * (MF * 8) + mf * 8 + g;
#pragma unroll
        for (int nf = 0; nf < NF; nf++) {
            const int col = n0 + wn * (NF * 8) + nf * 8 + 2 * t;
            double2* dst = reinterpret_cast<double2*>(C + (int64_t)row * ld + col);
            double2 v;
            if (cmode == C_STORE) {
                v.x = acc[mf][nf][0]; v.y = acc[mf][nf][1];
            } else if (cmode == C_NEG) {
                v.x = -acc[mf][nf][0]; v.y = -acc[mf][nf][1];
            } else {
                v = __ldcg(dst);
                v.x -= acc[mf][nf][0]; v.y -= acc[mf][nf][1];
            }
            *dst = v;
        }
    }
    __syncthreads();                                  // the ring is free for the next tile
}

// all tiles of one GEMM step that fall to this CTA: tile index (counted over non-skipped tiles) round-robin over the
// participating CTAs, starting at `first` so that consecutive independent steps interleave
template <int BL, int BM, int BN>
__device__ __forceinline__ void fused_gemm_step(const FusedOp& op, const double* A, const double* B, double* C, int64_t ld,
                                                int slot, int nslots, double* smem) {
    const int nI = op.M / BM, nJ = op.N / BN;
    int idx = 0;
    for (int it = 0; it < nI; it++) {
        // heaviest row tiles first for k <= i ranges
        const int I = (op.krange == KR_LE_I) ? nI - 1 - it : it;
        for (int jt = 0; jt < nJ; jt++) {
            const int J = (op.krange == KR_LE_J) ? nJ - 1 - jt : jt;
            const int m0 = I * BM, n0 = J * BN;
            if (op.lower_only && n0 > m0 + BM - 1) continue;
            if ((idx++ % nslots) != slot) continue;
            int kt_lo = 0, kt_hi = op.K / GEMM_BK;
            if (op.krange == KR_LE_J) kt_hi = min(kt_hi, (n0 + BN) / GEMM_BK);
            else if (op.krange == KR_GE_J) kt_lo = n0 / GEMM_BK;
            else if (op.krange == KR_LE_I) kt_hi = min(kt_hi, (m0 + BM) / GEMM_BK);
            else if (op.krange == KR_GE_I) kt_lo = m0 / GEMM_BK;
            fused_gemm_tile<BL, BM, BN>(A, B, C, ld, m0, n0, kt_lo, kt_hi, op.cmode, smem);
        }
    }
}

__global__ void __launch_bounds__(FUSED_THREADS, 1) fused_factor_kernel(const __grid_constant__ FusedArgs p) {
    extern __shared__ __align__(16) double fused_smem[];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int G = (int)cluster.num_blocks();
    const int b = blockIdx.y;
    double* mats[2] = {p.A + (int64_t)b * p.stride, p.Linv + (int64_t)b * p.stride};

    FUSED_STAMP(p.n_ops, 0);
    for (int i = 0; i < p.n_ops; i++) {
        const FusedOp& op = p.ops[i];
        if (op.type == FOP_LEAF) {
            if (rank == 0)
                leaf_body(mats[0] + (int64_t)op.r0 * p.ld + op.r0, mats[1] + (int64_t)op.r0 * p.ld + op.r0, p.ld, p.info + b,
                          op.r0, fused_smem);
        } else {
            int slot = rank, nslots = G;
            if (op.skip_rank0 && G > 1) { slot = rank - 1; nslots = G - 1; }
            if (slot >= 0) {
                const double* A = mats[op.a_mat] + (int64_t)op.a_off_r * p.ld + op.a_off_c;
                const double* B = mats[op.b_mat] + (int64_t)op.b_off_r * p.ld + op.b_off_c;
                double* C = mats[op.c_mat] + (int64_t)op.c_off_r * p.ld + op.c_off_c;
#define EXQ_FUSED_DISPATCH(BL)                                                                                     \
    switch (op.tile) {                                                                                             \
        case FT_32x32: fused_gemm_step<BL, 32, 32>(op, A, B, C, p.ld, slot, nslots, fused_smem); break;             \
        case FT_32x64: fused_gemm_step<BL, 32, 64>(op, A, B, C, p.ld, slot, nslots, fused_smem); break;             \
        case FT_64x128: fused_gemm_step<BL, 64, 128>(op, A, B, C, p.ld, slot, nslots, fused_smem); break;           \
        default: fused_gemm_step<BL, 128, 128>(op, A, B, C, p.ld, slot, nslots, fused_smem); break;                 \
    }
                if (op.b_layout == LAYOUT_K) { EXQ_FUSED_DISPATCH(LAYOUT_K) } else { EXQ_FUSED_DISPATCH(LAYOUT_MN) }
#undef EXQ_FUSED_DISPATCH
            }
        }
        FUSED_STAMP(i, 0);
        if (op.sync_after) {
            __threadfence();
            cluster.sync();
        }
        FUSED_STAMP(i, 1);
    }
}

// ---- host: the step list of a sub-tree ------------------------------------------------------------------------------
inline int fused_pick_tile(int M, int N, int lower_only, int G) {
    // the largest tile that still gives every CTA of the cluster something to do; else the smallest
    const int order[4] = {FT_128x128, FT_64x128, FT_32x64, FT_32x32};
    for (int k = 0; k < 4; k++) {
        const int bm = fused_tile_bm(order[k]), bn = fused_tile_bn(order[k]);
        if (M % bm || N % bn) continue;
        int tiles = 0;
        for (int i = 0; i < M / bm; i++)
            for (int j = 0; j < N / bn; j++)
                if (!(lower_only && j * bn > i * bm + bm - 1)) tiles++;
        if (tiles >= G) return order[k];
    }
    return FT_32x32;
}

// appends the steps of node(r0, n) to ops; returns false if the list would overflow
inline bool fused_build(FusedOp* ops, int& n_ops, int r0, int n, int G) {
    if (n == 128) {
        if (n_ops >= FUSED_MAX_OPS) return false;
        FusedOp o{};
        o.type = FOP_LEAF; o.r0 = r0; o.sync_after = 1;
        ops[n_ops++] = o;
        return true;
    }
    const int nb = n / 128;
    const int n1 = ((nb + 1) / 2) * 128, n2 = n - n1;
    if (!fused_build(ops, n_ops, r0, n1, G)) return false;
    if (n_ops + 4 > FUSED_MAX_OPS) return false;
    const int r1 = r0 + n1;
    auto gemm = [&](int am, int ar, int ac, int bm, int br, int bc, int bl, int cm, int cr, int cc, int M, int N, int K,
                    int krange, int lower, int cmode, int sync_after, int skip0) {
        FusedOp o{};
        o.type = FOP_GEMM;
        o.a_mat = am; o.a_off_r = ar; o.a_off_c = ac;
        o.b_mat = bm; o.b_off_r = br; o.b_off_c = bc; o.b_layout = bl;
        o.c_mat = cm; o.c_off_r = cr; o.c_off_c = cc;
        o.M = M; o.N = N; o.K = K; o.krange = krange; o.lower_only = lower; o.cmode = cmode;
        o.tile = fused_pick_tile(M, N, lower, skip0 ? (G > 1 ? G - 1 : 1) : G);
        o.sync_after = sync_after; o.skip_rank0 = skip0;
        ops[n_ops++] = o;
    };
    // (1) L21 = A21 Linv11^T  -> Linv21 slot
    gemm(0, r1, r0, 1, r0, r0, LAYOUT_K, 1, r1, r0, n2, n1, n1, KR_LE_J, 0, C_STORE, 1, 0);
    // (2) A22 -= L21 L21^T (lower tiles)
    gemm(1, r1, r0, 1, r1, r0, LAYOUT_K, 0, r1, r1, n2, n2, n1, KR_FULL, 1, C_SUB, 1, 0);
    // (3) T = L21 Linv11 -> A21 slot; independent of the leaf that opens node(A22): CTAs 1.. run it beside that leaf
    gemm(1, r1, r0, 1, r0, r0, LAYOUT_MN, 0, r1, r0, n2, n1, n1, KR_GE_J, 0, C_STORE, 0, 1);
    if (!fused_build(ops, n_ops, r1, n2, G)) return false;
    if (n_ops + 1 > FUSED_MAX_OPS) return false;
    // (4) Linv21 = -Linv22 T
    gemm(1, r1, r1, 0, r1, r0, LAYOUT_MN, 1, r1, r0, n2, n1, n2, KR_LE_I, 0, C_NEG, 1, 0);
    return true;
}

}  // namespace exq
