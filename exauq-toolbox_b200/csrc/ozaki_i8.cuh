// FP64-accurate variance product on the int8 tcgen05 tensor cores (Ozaki-style slicing).
//
// The predictive variance (mogp GaussianProcess.predict via exauq/core/emulators.py:649-667)
// needs  sumsq[c] = sum_i ( sum_{k<=i} Linv[i][k] * Kc[c][k] )^2  for every candidate c.  The
// FP64 tensor pipe (DMMA) tops out at ~36 TFLOP/s on this chip; tcgen05 has no f64 kind, but
// its int8 kind is exact (s32 accumulation in TMEM), so the product is rebuilt from integer
// pieces:
//
//   a = 2^e_row * sum_{p=1..S} a_p 256^-p,   a_p in [-128,127]   (balanced base-256 digits of the
//                                                                   value rounded to 8 S bits)
//   sum_k a_k b_k = 2^(e+f) * sum_{t=2..S+1} 256^-t  D_t,   D_t = sum_{p+q=t} sum_k a_p,k b_q,k
//
// All S(S+1)/2 slice products with p+q <= S+1 are issued as tcgen05.mma.kind::i8 into S
// accumulators (one per diagonal t) that stay in TMEM for the whole k loop; every D_t is an
// exact integer (|D_t| <= S 2^14 K < 2^31 for K <= 16384), the diagonals are combined exactly in
// int64 and rounded to FP64 once.  S = 6 (48 bits, 21 products) gives |error| ~ 2e-12 on the
// variance at N = 4096 (the FP64 path itself: ~1e-14; parity target 1e-8 relative), S = 7
// (56 bits, 28 products) is at the FP64 level (2e-14).
//
// Kernel layout (one CTA per SM, persistent over (candidate tile, row block) work items):
//   warp 0     TMA producer: per 64-byte k block one 3-D box {64 B, 128 cand, S} and one
//              {64 B, 64 rows, S} into a 2- or 3-stage shared-memory ring (SWIZZLE_64B, K-major)
//   warp 1     single-thread tcgen05.mma issuer (cta_group::1, M = 128, K = 32, N = 64 m: one digit of the
//              128-row operand against up to 4 consecutive digits of the 64-row operand per instruction),
//              commits to the stage's "empty" barrier and, after the last k block, to "tmem_full"
//   warps 2-9  epilogue: tcgen05.ld of the S diagonals, exact int64 recombination, row scaling,
//              per-candidate sum of squares over 32 of the 64 rows of the block -> P[2 rowblock + half][cand].
//              A warp may only read the TMEM lane quarter (warp % 4); two warps share a quarter and split the
//              64 columns, which halves the time the accumulators are held before the next item's MMAs may start.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "exq_common.cuh"

namespace exq {
namespace oz {

constexpr int W = 8;            // bits per slice (balanced base-256 digits)
constexpr int BLK_C = 128;      // candidates per tile  (UMMA M, TMEM lanes)
constexpr int BLK_R = 64;       // Linv rows per tile   (UMMA N, TMEM columns per diagonal)
constexpr int BLK_K = 64;       // k bytes per stage (one SWIZZLE_64B row)
constexpr int UMMA_K = 32;      // k per tcgen05.mma for 8-bit operands
constexpr int THREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter: 32 columns each)
constexpr int EPI_WARPS = 8;
constexpr int RASTER_G = 8;     // candidate tiles per L2 group

template <int S>
struct Cfg {
    static constexpr int A_BYTES = S * BLK_C * BLK_K;   // candidate slices of one stage
    static constexpr int B_BYTES = S * BLK_R * BLK_K;   // Linv slices of one stage
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (227 * 1024 - 1280) / STAGE_BYTES;   // 3 for S <= 6, 2 for S = 7
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
    static constexpr int TMEM_COLS = 512;
    static_assert(S * BLK_R <= 512, "accumulators exceed TMEM");
    static_assert(S >= 5 && S <= 7, "5 <= S <= 7 (8 S bits must fit an int64 with headroom)");
};

// exponent e with |v| * 2^-e < 0.49: the balanced digits then represent the value without a carry
// out of the leading digit
__host__ __device__ inline int scale_exponent(double mx) {
    if (!(mx > 0.0)) return 0;
    int e;
#ifdef __CUDA_ARCH__
    e = ilogb(mx) + 2;
    if (ldexp(mx, -e) >= 0.49) e++;
#else
    e = std::ilogb(mx) + 2;
    if (std::ldexp(mx, -e) >= 0.49) e++;
#endif
    // keep 2^(8S - e) and 2^(e - 40) normal numbers; rows below 2^-900 (or non-finite) degrade gracefully
    return e < -900 ? -900 : (e > 900 ? 900 : e);
}

__device__ __forceinline__ double pow2d(int e) {   // 2^e, normal range
    return __hiloint2double((1023 + e) << 20, 0);
}

// row exponents of a lower-triangular matrix (entries k <= row only)
__global__ void row_exponent_lower_kernel(const double* __restrict__ A, int64_t ld, int n, int* __restrict__ row_exp) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double mx = 0.0;
    for (int k = lane; k <= row; k += 32) mx = fmax(mx, fabs(A[(int64_t)row * ld + k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) row_exp[row] = scale_exponent(mx);
}

// doubles -> S int8 planes out[p][row][k]; row_exp == nullptr: one global exponent.
// lower != 0: entries with k > row are written as zeros (the source may hold scratch there).
template <int S>
__global__ void __launch_bounds__(256) slice_kernel(const double* __restrict__ src, int64_t ld, int rows, int cols,
                                                   const int* __restrict__ row_exp, int global_exp, int lower,
                                                   int8_t* __restrict__ out) {
    const int groups_per_row = cols >> 4;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)rows * groups_per_row) return;
    const int row = (int)(gid / groups_per_row);
    const int k0 = (int)(gid % groups_per_row) << 4;
    const int e = row_exp ? row_exp[row] : global_exp;
    const double sc = pow2d(8 * S - e);
    // adding 0x80 to every digit position turns the balanced digits into the plain bytes of x + BIAS
    constexpr unsigned long long BIAS = 0x8080808080808080ull >> (8 * (8 - S));
    uint32_t packed[S][4];
#pragma unroll
    for (int p = 0; p < S; p++) packed[p][0] = packed[p][1] = packed[p][2] = packed[p][3] = 0u;
    if (!lower || k0 <= row) {
        const double2* s2 = reinterpret_cast<const double2*>(src + (int64_t)row * ld + k0);
#pragma unroll
        for (int h = 0; h < 8; h++) {
            double2 v = s2[h];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int k = k0 + 2 * h + u;
                double r = (u ? v.y : v.x) * sc;
                if (lower && k > row) r = 0.0;
                const unsigned long long x = (unsigned long long)__double2ll_rn(r) + BIAS;
#pragma unroll
                for (int p = 0; p < S; p++) {
                    const uint32_t digit = ((uint32_t)(x >> (8 * (S - 1 - p))) & 0xffu) ^ 0x80u;   // plane 0 = leading digit
                    packed[p][(2 * h + u) >> 2] |= digit << (8 * ((2 * h + u) & 3));
                }
            }
        }
    }
    const int64_t plane = (int64_t)rows * cols;
#pragma unroll
    for (int p = 0; p < S; p++)
        *reinterpret_cast<uint4*>(out + p * plane + (int64_t)row * cols + k0) =
            make_uint4(packed[p][0], packed[p][1], packed[p][2], packed[p][3]);
}

// mbarrier wait that traps instead of hanging the GPU if a barrier is never completed
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t phase) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, phase); spins++)
        if (spins > (1u << 28)) __trap();
}

// ---- tcgen05 / TMEM primitives ----------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets lane (base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// K-major SWIZZLE_64B shared-memory matrix descriptor (rows of 64 B, 8-row groups 512 B apart)
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);     // start address
    d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;             // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                      // SWIZZLE_64B
    return d;
}
// instruction descriptor: s8 x s8 -> s32, both operands K-major, M x N
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Shared-memory rings of the two tcgen05 kernels.
//
// Round-2 layout (EXQ_OZ_K128 = 0, Cfg<S>): stages of 64 bytes of k holding all S digits of both operands, SWIZZLE_64B.  At
// S = 7 a stage is 86 KB, two fit, and ncu showed the tensor pipe 70 % active.  A ring of 32-byte stages (five fit) measured
// SLOWER, 69 against 86 TFLOP/s FP64-equivalent at n = 2048 (tools/micro/ozaki_gemm_test.cu), which pointed at the cause: the
// TMA unit pays per ROW of a box (~1.3 cycles per row per SM whether the row is 32 or 64 bytes), so a 64-byte stage of 1344
// rows needs ~1800 cycles to land plus the L2 round trip -- more than the ~2000 cycles its MMAs take when only two stages
// exist; and 64-byte rows use half the width of a shared-memory write.
//
// Current layout (EXQ_OZ_K128 = 1, KCfg<S>): k blocks of 128 bytes (SWIZZLE_128B, rows as wide as shared memory: half the rows
// per byte), and TWO rings so that they still fit: the S digit tiles of the 64-row operand of a k block form a stage
// (S x 8 KB, two stages), the digit tiles of the 128-row operand -- each is used by the MMAs of ONE digit p and then dead --
// stream through a ring of 16 KB slots (six to eight of them) with a barrier pair per slot.  Node products of the recursion
// 91 -> 97 TFLOP/s FP64-equivalent, LAUUM 71 -> 75 (n = 2048, 15 systems).
#ifndef EXQ_OZ_K128
#define EXQ_OZ_K128 1
#endif
constexpr int OZ_KB = EXQ_OZ_K128 ? 128 : BLK_K;     // k bytes per block of the pipeline
template <int S, int STATIC_SMEM = 0>
struct KCfg {
    // A slot = one digit tile of the 128-row operand; B stage = S digit tiles of the 64-row operand
    static constexpr int A_SLOT = BLK_C * 128;
    static constexpr int B_STAGE = S * BLK_R * 128;
    static constexpr int NB = 2;
    static constexpr int NA = (227 * 1024 - 1280 - STATIC_SMEM - NB * B_STAGE) / A_SLOT;
    static constexpr int K128_BYTES = NA * A_SLOT + NB * B_STAGE;
    static constexpr int SMEM_BYTES = (EXQ_OZ_K128 ? K128_BYTES : Cfg<S>::STAGES * Cfg<S>::STAGE_BYTES) + 1024 + 256;
    static constexpr int TMEM_COLS = 512;
    static_assert(NA >= 4, "A ring too short");
};
// The variance kernel takes the split rings only where the 64-byte layout is short of stages: S = 7 gains 4.5 % (4.85 ->
// 4.64 ms per 32768-candidate chunk at Np = 4096), S = 6 with its three 64-byte stages is 2 % faster as it was (3.55 ms).
template <int S>
__host__ __device__ constexpr bool var_k128() { return EXQ_OZ_K128 != 0 && S >= 7; }
template <int S>
__host__ __device__ constexpr int var_smem_bytes() { return var_k128<S>() ? KCfg<S>::K128_BYTES + 1024 + 256 : Cfg<S>::SMEM_BYTES; }
// K-major SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}

struct Args {
    const int* row_exp;   // exponent of every Linv row
    double* P;            // [2 Np / 64][ldp] column sums of squares per half row block (32 rows of Linv)
    int64_t ldp;
    int n_ctile;          // candidate tiles (Mc / 128)
    int n_rblk;           // row blocks (Np / 64)
    double out_scale;     // 2^(2 f - 80): global candidate-side scale and the diagonal weights
};

// work item i -> (candidate tile, row block): groups of RASTER_G candidate tiles, inside a group
// the heaviest row blocks (largest k range) first, candidate tiles of the group fastest
__device__ __forceinline__ void decode_item(int i, int n_ctile, int n_rblk, int& ct, int& j) {
    const int per_group = RASTER_G * n_rblk;
    const int gidx = i / per_group;
    const int r = i - gidx * per_group;
    const int g0 = gidx * RASTER_G;
    const int gsz = min(RASTER_G, n_ctile - g0);
    j = n_rblk - 1 - r / gsz;
    ct = g0 + r % gsz;
}
__device__ __forceinline__ int item_count(int n_ctile, int n_rblk) { return n_ctile * n_rblk; }

template <int S>
__global__ void __launch_bounds__(THREADS, 1)
colsumsq_kernel(const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_l, Args a) {
    using C = KCfg<S>;
    using C64 = Cfg<S>;
    constexpr bool K128 = var_k128<S>();
    constexpr int STAGES = K128 ? C::NB : C64::STAGES;      // stages of the (Linv-side or combined) ring
    constexpr int NA = K128 ? C::NA : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // K128: [NA candidate digit tiles][NB Linv stages][barriers]; else [STAGES combined stages][barriers]
    uint8_t* smem_b = smem + (K128 ? NA * C::A_SLOT : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(K128 ? smem_b + C::NB * C::B_STAGE : smem + STAGES * C64::STAGE_BYTES);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 1;
    uint64_t* full_a = bars + 2 * STAGES + 2;   // [NA]
    uint64_t* empty_a = full_a + NA;            // [NA]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty_a + NA);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = item_count(a.n_ctile, a.n_rblk);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < NA; s++) {
            mbar_init(&full_a[s], 1);
            mbar_init(&empty_a[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, EPI_WARPS);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int st = 0, sa_i = 0;
            uint32_t ph = 0, pha = 0;
            for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
                int ct, j;
                decode_item(i, a.n_ctile, a.n_rblk, ct, j);
                if (K128) {
                    // 128-byte k blocks 0 .. j/2 (for even j the second half lies above Linv's diagonal: zero digits)
                    for (int kb = 0; kb <= (j >> 1); kb++) {
                        mbar_wait_guard(&empty[st], ph ^ 1u);
                        mbar_expect_tx(&full[st], (uint32_t)C::B_STAGE);
                        tma_load_3d(smem_b + st * C::B_STAGE, &map_l, kb * 128, j * BLK_R, 0, &full[st]);
                        if (++st == STAGES) { st = 0; ph ^= 1u; }
#pragma unroll 1
                        for (int p = 0; p < S; p++) {
                            mbar_wait_guard(&empty_a[sa_i], pha ^ 1u);
                            mbar_expect_tx(&full_a[sa_i], (uint32_t)C::A_SLOT);
                            tma_load_3d(smem + sa_i * C::A_SLOT, &map_c, kb * 128, ct * BLK_C, p, &full_a[sa_i]);
                            if (++sa_i == NA) { sa_i = 0; pha ^= 1u; }
                        }
                    }
                } else {
                    for (int kb = 0; kb <= j; kb++) {
                        mbar_wait_guard(&empty[st], ph ^ 1u);
                        mbar_expect_tx(&full[st], (uint32_t)C64::STAGE_BYTES);
                        uint8_t* sa = smem + st * C64::STAGE_BYTES;
                        tma_load_3d(sa, &map_c, kb * BLK_K, ct * BLK_C, 0, &full[st]);
                        tma_load_3d(sa + C64::A_BYTES, &map_l, kb * BLK_K, j * BLK_R, 0, &full[st]);
                        if (++st == STAGES) { st = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int st = 0, sa_i = 0;
            uint32_t ph = 0, pha = 0, acc_ph = 0;
            for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
                int ct, j;
                decode_item(i, a.n_ctile, a.n_rblk, ct, j);
                mbar_wait_guard(tmem_empty, acc_ph ^ 1u);   // epilogue has drained the previous item's accumulators
                tc_fence_after();
                // digit p of the 128-row operand against digits q .. q+m-1 of the 64-row operand in ONE MMA of N = 64 m:
                // those digit tiles are contiguous in shared memory and their diagonals p+q .. p+q+m-1 are adjacent TMEM
                // column blocks, so the 128-row tile is read from shared memory once per 4 products instead of once per
                // product (at N = 64 the operand reads, 192 B/clk, exceed what shared memory delivers)
                if (K128) {
                    const int kb1 = (j >> 1) + 1;
                    for (int kb = 0; kb < kb1; kb++) {
                        mbar_wait_guard(&full[st], ph);
                        const uint64_t db0 = smem_desc_sw128(smem_u32(smem_b + st * C::B_STAGE));
#pragma unroll
                        for (int p = 0; p < S; p++) {
                            mbar_wait_guard(&full_a[sa_i], pha);
                            tc_fence_after();
                            const uint64_t da0 = smem_desc_sw128(smem_u32(smem + sa_i * C::A_SLOT));
#pragma unroll
                            for (int ks = 0; ks < 128 / UMMA_K; ks++) {
                                const uint64_t da = da0 + (uint64_t)((ks * UMMA_K) >> 4);
                                const uint32_t accumulate = (kb > 0 || ks > 0 || p > 0) ? 1u : 0u;
#pragma unroll
                                for (int q = 0; q + p < S; q += 4) {
                                    const int m = (S - p - q) < 4 ? (S - p - q) : 4;
                                    const uint64_t db = db0 + (uint64_t)((q * (BLK_R * 128) + ks * UMMA_K) >> 4);
                                    umma_i8(tmem_base + (uint32_t)((p + q) * BLK_R), da, db, idesc_i8(BLK_C, BLK_R * m), accumulate);
                                }
                            }
                            umma_commit(&empty_a[sa_i]);             // slot reusable once these MMAs have read it
                            if (++sa_i == NA) { sa_i = 0; pha ^= 1u; }
                        }
                        umma_commit(&empty[st]);
                        if (kb == kb1 - 1) umma_commit(tmem_full);   // accumulators complete
                        if (++st == STAGES) { st = 0; ph ^= 1u; }
                    }
                } else {
                    for (int kb = 0; kb <= j; kb++) {
                        mbar_wait_guard(&full[st], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + st * C64::STAGE_BYTES);
                        const uint64_t da0 = smem_desc_sw64(sa);
                        const uint64_t db0 = smem_desc_sw64(sa + C64::A_BYTES);
#pragma unroll
                        for (int ks = 0; ks < BLK_K / UMMA_K; ks++) {
#pragma unroll
                            for (int p = 0; p < S; p++) {
                                const uint64_t da = da0 + (uint64_t)((p * (BLK_C * BLK_K) + ks * UMMA_K) >> 4);
                                const uint32_t accumulate = (kb > 0 || ks > 0 || p > 0) ? 1u : 0u;
#pragma unroll
                                for (int q = 0; q + p < S; q += 4) {
                                    const int m = (S - p - q) < 4 ? (S - p - q) : 4;
                                    const uint64_t db = db0 + (uint64_t)((q * (BLK_R * BLK_K) + ks * UMMA_K) >> 4);
                                    umma_i8(tmem_base + (uint32_t)((p + q) * BLK_R), da, db, idesc_i8(BLK_C, BLK_R * m), accumulate);
                                }
                            }
                        }
                        umma_commit(&empty[st]);                 // stage reusable once these MMAs have read it
                        if (kb == j) umma_commit(tmem_full);     // accumulators complete
                        if (++st == STAGES) { st = 0; ph ^= 1u; }
                    }
                }
                acc_ph ^= 1u;
            }
        }
    } else {
        // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t acc_ph = 0;
        constexpr double LO_W = 1.0 / (double)(1ull << (W * (S - 4)));   // weight of the low group relative to the high one
        for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
            int ct, j;
            decode_item(i, a.n_ctile, a.n_rblk, ct, j);
            mbar_wait_guard(tmem_full, acc_ph);
            tc_fence_after();
            // phase 1: drain this warp's 32 columns of every diagonal into FP64 (exact int64 recombination), then hand
            // TMEM back at once -- the scaling and the sum of squares below overlap the next item's MMAs
            double z[32];
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                int32_t d[S][16];
#pragma unroll
                for (int u = 0; u < S; u++) tmem_ld16(tmem_base + lane_base + (uint32_t)(u * BLK_R + half * 32 + c0), d[u]);
                tmem_ld_wait();
                if (c0 == 16) {
                    // the last raw values are in registers: TMEM goes back BEFORE they are recombined
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty);
                }
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    // diagonals u = 0..3 (weights 256^-2 .. 256^-5) and u = 4..S-1, each combined exactly
                    int64_t hi = 0, lo = 0;
#pragma unroll
                    for (int u = 0; u < 4 && u < S; u++) hi = hi * (1 << W) + (int64_t)d[u][c];
#pragma unroll
                    for (int u = 4; u < S; u++) lo = lo * (1 << W) + (int64_t)d[u][c];
                    double v = (double)hi;
                    if (S > 4) v = fma((double)lo, LO_W, v);
                    z[c0 + c] = v;
                }
            }
            // phase 2
            double sum = 0.0;
            const int* re = a.row_exp + j * BLK_R + half * 32;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const double zz = z[c] * pow2d(__ldg(re + c));
                sum = fma(zz, zz, sum);
            }
            a.P[(int64_t)(2 * j + half) * a.ldp + (int64_t)ct * BLK_C + quarter * 32 + lane] = sum * a.out_scale;
            acc_ph ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// planes[S][rows][cols] int8 -> 3-D map {cols, rows, S}, box {64, box_rows, S}, SWIZZLE_64B
// box_k: k bytes per box = swizzle span (128 / 64 / 32); box_planes: digit planes per box (default: all S)
inline int make_plane_map(CUtensorMap* map, const int8_t* planes, int64_t rows, int64_t cols, int S, int box_rows,
                          int box_k = BLK_K, int box_planes = 0) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return -1;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)cols, (cuuint64_t)(rows * cols)};
    cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, (cuuint32_t)(box_planes > 0 ? box_planes : S)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    box_k == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : box_k == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

template <int S>
inline cudaError_t set_attrs() {
    return cudaFuncSetAttribute(colsumsq_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, var_smem_bytes<S>());
}

// P[2 j + h][c] = sum over rows 32 h .. 32 h + 31 of block j of (Linv k_c)^2, from sliced operands.
// cand_planes: [S][Mc][Np] (global exponent cand_exp), linv_planes: [S][Np][Np] with row_exp.
template <int S>
inline int launch_colsumsq(const int8_t* cand_planes, int Mc, int cand_exp, const int8_t* linv_planes, const int* row_exp,
                           int Np, double* P, int64_t ldp, int sm_count, cudaStream_t stream) {
    CUtensorMap mc, ml;
    int rc;
    constexpr bool K128 = var_k128<S>();
    if (K128 && (Np % 128)) return -2;
    // K128: one candidate digit tile per box, all S digit tiles of the Linv rows per box
    if ((rc = make_plane_map(&mc, cand_planes, Mc, Np, S, BLK_C, K128 ? 128 : BLK_K, K128 ? 1 : S))) return rc;
    if ((rc = make_plane_map(&ml, linv_planes, Np, Np, S, BLK_R, K128 ? 128 : BLK_K, S))) return rc;
    Args a{};
    a.row_exp = row_exp; a.P = P; a.ldp = ldp; a.n_ctile = Mc / BLK_C; a.n_rblk = Np / BLK_R;
    a.out_scale = std::ldexp(1.0, 2 * cand_exp - 2 * W * 5);   // (2^f 256^-5)^2
    const int items = a.n_ctile * a.n_rblk;
    const int grid = items < sm_count ? items : sm_count;
    colsumsq_kernel<S><<<grid, THREADS, var_smem_bytes<S>(), stream>>>(mc, ml, a);
    return 0;
}

template <int S>
inline void launch_slice(const double* src, int64_t ld, int rows, int cols, const int* row_exp, int global_exp, int lower,
                         int8_t* out, cudaStream_t stream) {
    const int64_t groups = (int64_t)rows * (cols >> 4);
    slice_kernel<S><<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(src, ld, rows, cols, row_exp, global_exp, lower, out);
}

}  // namespace oz
}  // namespace exq
