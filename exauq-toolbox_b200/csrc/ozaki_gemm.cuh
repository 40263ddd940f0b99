// General FP64-accurate tile GEMM on the int8 tcgen05 tensor cores, for the large nodes of the
// potrf + trtri recursion and for LAUUM (K^-1 = Linv^T Linv) inside the batched log-posterior
// evaluations (mogp GaussianProcess.logposterior / logpost_deriv as driven by
// exauq/utilities/mogp_fitting.py:289-346).  Same digit scheme, pipeline and TMEM layout as
// ozaki_i8.cuh; what differs is
//   * both operands are intermediate FP64 results, so every GEMM is preceded by an exponent
//     pass and a slicing pass over each operand (O(n^2) against the O(n^3) product); operands
//     stored k-strided (B of the triangular-inverse merge, both operands of LAUUM) are
//     transposed by the slicer, so the kernel only ever sees k-contiguous digit planes;
//   * the epilogue writes FP64 tiles  C (op)= 2^(ea_i + eb_j) * sum_t 256^-t D_t.
//
//   C[i][j] (op)= sum_k A(i, k) * B(j, k)     i < M, j < N, k in a per-tile range (KR_*)
#pragma once
#include <algorithm>
#include <vector>

#include "gemm_dmma.cuh"   // KR_*, C_* enums
#include "ozaki_i8.cuh"

namespace exq {
namespace oz {

#ifndef EXQ_OZ_SLICE_REG
#define EXQ_OZ_SLICE_REG 1      // register-resident straight slicer for rows of 512 / 1024 / 2048 entries
#endif

enum : int { MASK_NONE = 0, MASK_LE = 1 /* keep k <= row */, MASK_GE = 2 /* keep k >= row */ };

__device__ __forceinline__ bool mask_keep(int mask, int row, int k) {
    return mask == MASK_NONE || (mask == MASK_LE ? k <= row : k >= row);
}

// exponent of every operand row; the operand is src[row][k] (trans = 0) or src[k][row] (trans = 1)
// one warp per row (trans = 0); for trans = 1 a 32 x 8 thread block scans 32 columns
__global__ void __launch_bounds__(256) operand_exponent_kernel(const double* __restrict__ src, int64_t ld, int64_t stride_sys,
                                                               int rows, int K, int trans, int mask,
                                                               int* __restrict__ row_exp) {
    const double* s = src + (int64_t)blockIdx.z * stride_sys;
    int* out = row_exp + (int64_t)blockIdx.z * rows;
    if (!trans) {
        const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
        if (row >= rows) return;
        const int k_lo = mask == MASK_GE ? row : 0, k_hi = mask == MASK_LE ? min(row + 1, K) : K;
        double mx = 0.0;
        for (int k = k_lo + lane; k < k_hi; k += 32) mx = fmax(mx, fabs(s[(int64_t)row * ld + k]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) out[row] = scale_exponent(mx);
    } else {
        __shared__ double red[8][33];
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        const int row = blockIdx.x * 32 + tx;   // operand row = source column
        double mx = 0.0;
        if (row < rows) {
            const int k_lo = mask == MASK_GE ? row : 0, k_hi = mask == MASK_LE ? min(row + 1, K) : K;
            for (int k = k_lo + ty; k < k_hi; k += 8) mx = fmax(mx, fabs(s[(int64_t)k * ld + row]));
        }
        red[ty][tx] = mx;
        __syncthreads();
        if (ty == 0 && row < rows) {
#pragma unroll
            for (int q = 1; q < 8; q++) mx = fmax(mx, red[q][tx]);
            out[row] = scale_exponent(mx);
        }
    }
}

// exponents of the COLUMNS of an inverse factor from its column sums of squares (dinv_j = sum_k Linv[k][j]^2, which
// the triangular mat-vec pass has already produced): max_k |Linv[k][j]| <= sqrt(dinv_j), and >= sqrt(dinv_j / n), so
// the exponent is at most log2(sqrt n) = 6 bits (n = 4096) looser than the exact one -- usually 1-2, the diagonal entry
// carries most of a column's mass.  Saves the separate exponent pass over the whole matrix in front of LAUUM.
__global__ void __launch_bounds__(256) exponent_from_sumsq_kernel(const double* __restrict__ sumsq, int64_t stride, int n,
                                                                  int* __restrict__ row_exp) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double bound = sqrt(sumsq[(int64_t)blockIdx.z * stride + j]) * 1.0000001;
    row_exp[(int64_t)blockIdx.z * n + j] = scale_exponent(bound);
}

// exponents of the COLUMNS of a product from the column maxima its epilogue left behind (GArgs::colmax: the high word of
// max |value| per column, collected with redux + atomicMax while the tiles were written): the bound is the largest double
// with that high word, so the exponent is exact except when max |value| sits within 2^-20 of a rounding boundary of
// scale_exponent (then one larger: one bit of one column).  Saves the pass over the matrix that operand_exponent_kernel makes.
__global__ void __launch_bounds__(256) exponent_from_colmax_kernel(const uint32_t* __restrict__ colmax, int64_t stride, int n,
                                                                   int* __restrict__ row_exp) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t hi = colmax[(int64_t)blockIdx.z * stride + j];
    const double bound = hi ? __hiloint2double((int)hi, (int)0xffffffffu) : 0.0;
    row_exp[(int64_t)blockIdx.z * n + j] = scale_exponent(bound);
}

// digits of S planes from a double scaled by 2^(8S - e)
template <int S>
__device__ __forceinline__ unsigned long long digits_of(double v, double sc) {
    constexpr unsigned long long BIAS = 0x8080808080808080ull >> (8 * (8 - S));
    return (unsigned long long)__double2ll_rn(v * sc) + BIAS;   // byte (S-1-p) ^ 0x80 = digit of plane p
}

// straight slicer for a batch of operands: out[p][b * rows + row][k]  (planes of all systems stacked).
// One warp per row: the row maximum fixes the exponent, then the same warp re-reads the row (L1/L2
// resident) and writes its digits, so the operand crosses HBM once.
template <int S>
__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ src, int64_t ld, int64_t stride_sys,
                                                         int rows, int K, int mask, int* __restrict__ row_exp,
                                                         int8_t* __restrict__ out, int64_t plane_bytes) {
    const int b = blockIdx.z, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double* srow = src + (int64_t)b * stride_sys + (int64_t)row * ld;
    // 16-element groups that hold kept entries
    const int g_lo = mask == MASK_GE ? row >> 4 : 0, g_hi = mask == MASK_LE ? (row >> 4) + 1 : K >> 4;
    double mx = 0.0;
    for (int k0 = (g_lo << 4) + (lane << 1); k0 < (g_hi << 4); k0 += 64) {
        const double2 v = *reinterpret_cast<const double2*>(srow + k0);
        if (mask_keep(mask, row, k0)) mx = fmax(mx, fabs(v.x));
        if (mask_keep(mask, row, k0 + 1)) mx = fmax(mx, fabs(v.y));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int e = scale_exponent(mx);
    if (lane == 0) row_exp[(int64_t)b * rows + row] = e;
    const double sc = pow2d(8 * S - e);
    int8_t* orow = out + ((int64_t)b * rows + row) * K;
    // The kernels read a masked operand only inside the (128-byte aligned) k range of the tile the row belongs to: zero
    // digits are needed from the diagonal to the end of its 128-block (MASK_LE) or from the start of it (MASK_GE), not over
    // the whole row -- the rest of the planes is never read (tile_krange) and is left as it is.
    const int w_lo = mask == MASK_GE ? (row & ~127) : 0, w_hi = mask == MASK_LE ? min(K, (row & ~127) + 128) : K;
    // 4 consecutive elements per lane: 1 KB contiguous per warp on the way in, 128 B per plane on the way out
    for (int k0 = w_lo + (lane << 2); k0 < w_hi; k0 += 128) {
        uint32_t packed[S];
#pragma unroll
        for (int p = 0; p < S; p++) packed[p] = 0u;
        if ((k0 >> 4) >= g_lo && (k0 >> 4) < g_hi) {
            const double2* s2 = reinterpret_cast<const double2*>(srow + k0);
            const double2 v0 = s2[0], v1 = s2[1];
            const double vals[4] = {v0.x, v0.y, v1.x, v1.y};
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double r = mask_keep(mask, row, k0 + u) ? vals[u] : 0.0;
                const unsigned long long x = digits_of<S>(r, sc);
#pragma unroll
                for (int p = 0; p < S; p++) packed[p] |= (((uint32_t)(x >> (8 * (S - 1 - p))) & 0xffu) ^ 0x80u) << (8 * u);
            }
        }
#pragma unroll
        for (int p = 0; p < S; p++) *reinterpret_cast<uint32_t*>(orow + p * plane_bytes + k0) = packed[p];
    }
}

// Straight slicer for rows of 512 / 1024 / 2048 entries (the half sizes of the recursion's nodes): the warp's row is held in
// registers (4 consecutive entries per lane and 128-entry step: TRIPS x 4 doubles per lane), so every load of the row is in
// flight before the first is used and the row is read once -- slice_rows_kernel re-reads it through L1 after the maximum
// is known and keeps two loads in flight per lane (0.66 of the HBM rate at n = 2048).  Same digits, same exponents.
template <int S, int TRIPS>
__global__ void __launch_bounds__(256) slice_rows_reg_kernel(const double* __restrict__ src, int64_t ld, int64_t stride_sys,
                                                             int rows, int mask, int* __restrict__ row_exp,
                                                             int8_t* __restrict__ out, int64_t plane_bytes) {
    constexpr int K = TRIPS * 128;
    const int b = blockIdx.z, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double* srow = src + (int64_t)b * stride_sys + (int64_t)row * ld;
    // 16-element groups that hold kept entries (as slice_rows_kernel)
    const int g_lo = mask == MASK_GE ? row >> 4 : 0, g_hi = mask == MASK_LE ? (row >> 4) + 1 : K >> 4;
    double v[TRIPS][4];
#pragma unroll
    for (int i = 0; i < TRIPS; i++) {
        const int k0 = (lane << 2) + 128 * i;
        const bool in = (k0 >> 4) >= g_lo && (k0 >> 4) < g_hi;
        double2 a = make_double2(0.0, 0.0), c = make_double2(0.0, 0.0);
        if (in) {
            const double2* s2 = reinterpret_cast<const double2*>(srow + k0);
            a = s2[0];
            c = s2[1];
        }
        v[i][0] = mask_keep(mask, row, k0) ? a.x : 0.0;
        v[i][1] = mask_keep(mask, row, k0 + 1) ? a.y : 0.0;
        v[i][2] = mask_keep(mask, row, k0 + 2) ? c.x : 0.0;
        v[i][3] = mask_keep(mask, row, k0 + 3) ? c.y : 0.0;
    }
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < TRIPS; i++)
#pragma unroll
        for (int u = 0; u < 4; u++) mx = fmax(mx, fabs(v[i][u]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int e = scale_exponent(mx);
    if (lane == 0) row_exp[(int64_t)b * rows + row] = e;
    const double sc = pow2d(8 * S - e);
    int8_t* orow = out + ((int64_t)b * rows + row) * K;
    // zero digits only up to the end of / from the start of the 128-byte block of the diagonal (see slice_rows_kernel)
    const int w_lo = mask == MASK_GE ? (row & ~127) : 0, w_hi = mask == MASK_LE ? min(K, (row & ~127) + 128) : K;
#pragma unroll
    for (int i = 0; i < TRIPS; i++) {
        const int k0 = (lane << 2) + 128 * i;
        if (k0 < w_lo || k0 >= w_hi) continue;
        uint32_t packed[S];
#pragma unroll
        for (int p = 0; p < S; p++) packed[p] = 0u;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned long long x = digits_of<S>(v[i][u], sc);
#pragma unroll
            for (int p = 0; p < S; p++) packed[p] |= (((uint32_t)(x >> (8 * (S - 1 - p))) & 0xffu) ^ 0x80u) << (8 * u);
        }
#pragma unroll
        for (int p = 0; p < S; p++) *reinterpret_cast<uint32_t*>(orow + p * plane_bytes + k0) = packed[p];
    }
}

// transposing slicer: operand row r = source column, k = source row; 64 x 64 tiles through shared memory
template <int S>
__global__ void __launch_bounds__(256) slice_cols_kernel(const double* __restrict__ src, int64_t ld, int64_t stride_sys,
                                                         int rows, int K, int mask, const int* __restrict__ row_exp,
                                                         int8_t* __restrict__ out, int64_t plane_bytes) {
    __shared__ uint32_t dig[S][64][17];   // [plane][operand row][4 k per word], 17-word rows: conflict-free column writes
    const int b = blockIdx.z, r0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int rr = threadIdx.x & 63, kq = threadIdx.x >> 6;
    const int row = r0 + rr;
    // tiles entirely outside the kept triangle are all zeros
    const bool any = mask == MASK_NONE || (mask == MASK_LE ? k0 <= r0 + 63 : k0 + 63 >= r0);
    // tiles outside the kept triangle: zero digits only inside the 128-block of the diagonal, nothing beyond it is ever
    // read (see slice_rows_kernel)
    if (!any && (mask == MASK_LE ? k0 >= (r0 & ~127) + 128 : k0 + 64 <= (r0 & ~127))) return;
    const double sc = pow2d(8 * S - row_exp[(int64_t)b * rows + row]);
    const double* s = src + (int64_t)b * stride_sys;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int w = kq + 4 * i;           // word = 4 consecutive k
        uint32_t packed[S];
#pragma unroll
        for (int p = 0; p < S; p++) packed[p] = 0u;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = k0 + 4 * w + u;
            double v = 0.0;
            if (any && mask_keep(mask, row, k)) v = s[(int64_t)k * ld + row];
            const unsigned long long x = digits_of<S>(v, sc);
#pragma unroll
            for (int p = 0; p < S; p++) packed[p] |= (((uint32_t)(x >> (8 * (S - 1 - p))) & 0xffu) ^ 0x80u) << (8 * u);
        }
#pragma unroll
        for (int p = 0; p < S; p++) dig[p][rr][w] = packed[p];
    }
    __syncthreads();
    // 64 rows x 64 bytes per plane = 256 x 16 B per plane: one per thread
    const int orow = threadIdx.x >> 2, oq = threadIdx.x & 3;
#pragma unroll
    for (int p = 0; p < S; p++) {
        const uint4 v = make_uint4(dig[p][orow][oq * 4], dig[p][orow][oq * 4 + 1], dig[p][orow][oq * 4 + 2], dig[p][orow][oq * 4 + 3]);
        *reinterpret_cast<uint4*>(out + p * plane_bytes + ((int64_t)b * rows + r0 + orow) * K + k0 + oq * 16) = v;
    }
}

// cmode == C_GRAD (LAUUM only, kernels instantiated with GK >= 0): the tile of K^-1 = Linv^T Linv is not stored; the
// epilogue contracts it with dK/dtheta on the spot --
//     grad_k = 1/2 sum_ij (K^-1 - alpha alpha^T)_ij dK_ij/d raw_k     (mogp logpost_deriv, exauq/utilities/mogp_fitting.py:318-324)
// -- while the tensor pipe runs the next tile: the ~65 FP64 operations per entry that used to be a kernel of their own
// (grad_fast_kernel: 1.45 ms per round of 15 systems, FP64 pipe 35 % active at 8 warps per SM) ride on the idle
// FP64 pipe under the int8 MMAs, and K^-1 (1 GB per round) never goes to HBM.  Same pair arithmetic as
// grad_fast_kernel; per (system, CTA, epilogue warp) partial sums in fixed order, reduced by grad_reduce_slots_kernel.
constexpr int GRAD_D = 8;       // input dimensions the fused epilogue holds in registers (d <= 8)
#ifndef EXQ_GRAD_ILP
#define EXQ_GRAD_ILP 4
#endif
constexpr int GRAD_ILP = EXQ_GRAD_ILP;   // (i, j) pairs in flight per epilogue thread
struct GradArgs {
    const double* X;        // [Np][d] training inputs, zero padded (shared by all systems)
    const double* theta;    // [B][d + 2]: exp(raw_k), sigma2, nugget
    const double* alpha;    // [B][ldv]
    int64_t ldv;
    int n_valid, d;
    double* partial;        // [B][slots][GRAD_D + 2], zeroed by the host; slot = CTA * EPI_WARPS + epilogue warp
    int slots;
};

struct GArgs {
    GradArgs g;
    const int* expA;      // [B][M]
    const int* expB;      // [B][N]
    double* C;
    int64_t ldc, strideC;
    int M, N, K;          // multiples of 128 / 64 / 64
    int B;
    int krange, lower_only, cmode;
    int heavy_last;       // tiles with larger I (or J) carry more k blocks: walk the tile list backwards
    int sys_slow;         // 1: systems are the slow index of the work list (see decode_gemm_item)
    const uint32_t* sched;   // balanced work list (make_schedule): entry i is the item of CTA i % grid in round i / grid,
    int n_sched;             // packed (system << 20 | I << 10 | J), SCHED_NONE = idle; nullptr: the arithmetic walk above
    int sched_grid;          // CTAs the list was dealt to
    uint32_t* colmax;        // C_STORE only, may be nullptr: [B][colmax_stride] high words of max |C[:, j]| per column j (atomicMax,
    int64_t colmax_stride;   // zeroed by the host): the next product slices C transposed and needs its column exponents
};
constexpr uint32_t SCHED_NONE = 0xffffffffu;

// k-block range [lo, hi) of output tile (I: 128 rows, J: 64 columns)
__host__ __device__ __forceinline__ void tile_krange(int krange, int I, int J, int nkb, int& lo, int& hi) {
    lo = 0; hi = nkb;
    if (krange == KR_LE_J) hi = nkb < J + 1 ? nkb : J + 1;
    else if (krange == KR_GE_J) lo = J;
    else if (krange == KR_LE_I) hi = nkb < 2 * I + 2 ? nkb : 2 * I + 2;
    else if (krange == KR_GE_I) lo = 2 * I;
}
__host__ __device__ __forceinline__ bool tile_skipped(int lower_only, int I, int J) {
    return lower_only && J * BLK_R > I * BLK_C + (BLK_C - 1);
}
// k blocks (64 bytes of k for a 128 x 64 output tile) the kernel walks for one system: the SAME tile walk as
// decode_gemm_item / tile_krange, so that the int8 operation count of the roofline is what was issued
inline int64_t gemm_kblocks(int M, int N, int K, int krange, int lower_only) {
    const int nI = M / BLK_C, nJ = N / BLK_R, nkb = K / BLK_K;
    int64_t total = 0;
    for (int I = 0; I < nI; I++)
        for (int J = 0; J < nJ; J++) {
            if (tile_skipped(lower_only, I, J)) continue;
            int lo, hi;
            tile_krange(krange, I, J, nkb, lo, hi);
            if (hi <= lo) continue;
            // the kernel walks whole 128-byte blocks: ranges are widened to even 64-byte boundaries
            total += EXQ_OZ_K128 ? 2 * (((hi + 1) >> 1) - (lo >> 1)) : hi - lo;
        }
    return total;
}
// int8 operations (2 per multiply-accumulate) issued for B systems with S digits: S(S+1)/2 digit products per k block
inline double gemm_issued_ops(int M, int N, int K, int B, int krange, int lower_only, int S) {
    return 2.0 * BLK_C * BLK_R * BLK_K * (double)gemm_kblocks(M, N, K, krange, lower_only) * B * (S * (S + 1) / 2);
}
// linear item -> (system, I, J); returns false for tiles skipped by lower_only.
// sys_slow = 1 (full products): systems are the SLOW index, so the CTAs running at any moment work on one
// system (two at a boundary) whose digit planes (2 x 29 MB at n = k = 2048, S = 7) stay in L2 while its tiles
// reuse them; with the system as the fast index every CTA streamed its operand tiles from DRAM (ncu: 5.1 GB
// read per launch against 0.9 GB of planes, 45 % DRAM utilisation) -- measured 81 -> 89..96 TFLOP/s
// FP64-equivalent.  The lower-only products (SYRK, LAUUM) measure 7..12 % slower that way at n = 2048 (their
// skipped upper tiles unbalance the static round-robin inside a system) but LAUUM at n = 4096, whose planes
// exceed L2 per system, gains more than they lose: the design batch is fastest with sys_slow = 1 everywhere.
__device__ __forceinline__ bool decode_gemm_item(const GArgs& a, int item, int& b, int& I, int& J) {
    const int nJ = a.N / BLK_R, nI = a.M / BLK_C;
    const int tiles = nI * nJ;
    int t;
    if (a.sys_slow) {
        b = item / tiles;
        t = item - b * tiles;
    } else {
        t = item / a.B;
        b = item - t * a.B;
    }
    if (a.heavy_last) t = tiles - 1 - t;
    I = t / nJ;
    J = t - I * nJ;
    return !tile_skipped(a.lower_only, I, J);
}

// item i of this launch: from the balanced list when there is one
__device__ __forceinline__ bool gemm_item(const GArgs& a, int i, int& b, int& I, int& J) {
    if (a.sched) {
        const uint32_t it = __ldg(a.sched + i);
        if (it == SCHED_NONE) return false;
        b = (int)(it >> 20); I = (int)((it >> 10) & 1023u); J = (int)(it & 1023u);
        return b < a.B;       // the list may have been built for a larger batch: any prefix of it is balanced too
    }
    return decode_gemm_item(a, i, b, I, J);
}

// Balanced static schedule.  The arithmetic walk hands item i to CTA i % grid whatever the item weighs: tiles skipped by
// lower_only are holes in the list and the k range of a tile varies 1 : 32 across a triangular product, so the busiest CTA
// carried 7 % (L21, T), 23 % (SYRK) and 6-38 % (LAUUM) more k blocks than the average (15 systems, n = 2048 / 4096).  Here the
// kept tiles of each system are sorted by weight, systems stay in sequence (their digit planes share L2), and every round
// of `grid` items is dealt heaviest-first to the CTAs with the least work so far: the loads end within ~1 % of each other.
// Deterministic (a pure function of the shape), so results stay reproducible run to run.  Because the dealing is incremental,
// every prefix of the list that ends at a round boundary is balanced as well: a list built for B systems serves any smaller
// batch (entries of systems >= B are skipped by the kernel).  Returns the kept tiles per system.
inline int make_schedule(int M, int N, int K, int B, int krange, int lower_only, int grid, std::vector<uint32_t>& out) {
    const int nI = M / BLK_C, nJ = N / BLK_R, nkb = K / BLK_K;
    struct It { uint32_t id; float w; };
    std::vector<It> all;
    all.reserve((size_t)nI * nJ * B);
    std::vector<It> sys;
    for (int b = 0; b < B; b++) {
        sys.clear();
        for (int I = 0; I < nI; I++)
            for (int J = 0; J < nJ; J++) {
                if (tile_skipped(lower_only, I, J)) continue;
                int lo, hi;
                tile_krange(krange, I, J, nkb, lo, hi);
                if (hi <= lo) { lo = 0; hi = 0; }
                // 128-byte k blocks plus about one block's worth of fill + epilogue hand-over per tile
                const float w = (float)(EXQ_OZ_K128 ? ((hi + 1) >> 1) - (lo >> 1) : (hi - lo) * 0.5f) + 1.0f;
                sys.push_back({((uint32_t)b << 20) | ((uint32_t)I << 10) | (uint32_t)J, w});
            }
        std::stable_sort(sys.begin(), sys.end(), [](const It& x, const It& y) { return x.w > y.w; });
        all.insert(all.end(), sys.begin(), sys.end());
    }
    const size_t rounds = (all.size() + grid - 1) / grid;
    out.assign(rounds * grid, SCHED_NONE);
    std::vector<float> load(grid, 0.0f);
    std::vector<int> order(grid);
    std::vector<It> win;
    for (size_t r = 0; r < rounds; r++) {
        const size_t lo = r * grid, hi = std::min(all.size(), lo + grid);
        win.assign(all.begin() + lo, all.begin() + hi);
        std::stable_sort(win.begin(), win.end(), [](const It& x, const It& y) { return x.w > y.w; });
        for (int c = 0; c < grid; c++) order[c] = c;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return load[x] < load[y]; });
        for (size_t k = 0; k < win.size(); k++) {
            out[lo + order[k]] = win[k].id;
            load[order[k]] += win[k].w;
        }
    }
    return (int)(all.size() / (size_t)std::max(B, 1));
}

constexpr int GEMM_STATIC_SMEM = 6 * 1024;                   // gx_s / ga_s of the gradient epilogue + slack
template <int S>
using GCfg = KCfg<S, GEMM_STATIC_SMEM>;

// GK: -1 plain GEMM; EXQ_KERNEL_MATERN52 / EXQ_KERNEL_SQEXP: cmode C_GRAD is available (fused gradient contraction)
// (320 threads are budgeted as 384 by the register allocator: 168 registers per thread is the hardware limit here --
// a launch with 192 fails with "too many resources" -- so the gradient epilogue lives with ~200 bytes of spills)
template <int S, int GK = -1>
__global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GArgs a) {
    using C = GCfg<S>;
    using C64 = Cfg<S>;
    constexpr bool K128 = EXQ_OZ_K128 != 0;
    constexpr int STAGES = K128 ? C::NB : C64::STAGES;      // stages of the (B or combined) ring
    constexpr int NA = K128 ? C::NA : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // K128: [NA A slots][NB B stages][barriers]; else [STAGES combined stages][barriers]
    uint8_t* smem_b = smem + (K128 ? NA * C::A_SLOT : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(K128 ? smem_b + C::NB * C::B_STAGE : smem + STAGES * C64::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 1;
    uint64_t* full_a = bars + 2 * STAGES + 2;                // [NA]
    uint64_t* empty_a = full_a + NA;                         // [NA]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty_a + NA);

    // fused gradient epilogue: the tile's 64 columns (pre-scaled coordinates, alpha), shared by the epilogue warps
    __shared__ __align__(16) double gx_s[(GK >= 0 ? GRAD_D : 2) * (GK >= 0 ? BLK_R : 1)];
    __shared__ double ga_s[GK >= 0 ? BLK_R : 1];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.sched ? a.n_sched : (a.M / BLK_C) * (a.N / BLK_R) * a.B;
    const int nkb = a.K / BLK_K;

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
        if (lane == 0) {
            int st = 0, sa_i = 0;
            uint32_t ph = 0, pha = 0;
            for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
                int b, I, J, lo, hi;
                if (!gemm_item(a, i, b, I, J)) continue;
                tile_krange(a.krange, I, J, nkb, lo, hi);
                if (K128) {
                    // 128-byte k blocks: the range is widened to whole blocks (the extra 64 bytes lie beyond the operand's
                    // triangle, where the slicers wrote zero digits)
                    for (int kb = lo >> 1; kb < (hi + 1) >> 1; kb++) {
                        mbar_wait_guard(&empty[st], ph ^ 1u);
                        mbar_expect_tx(&full[st], (uint32_t)C::B_STAGE);
                        tma_load_3d(smem_b + st * C::B_STAGE, &map_b, kb * 128, b * a.N + J * BLK_R, 0, &full[st]);
                        if (++st == STAGES) { st = 0; ph ^= 1u; }
#pragma unroll 1
                        for (int p = 0; p < S; p++) {
                            mbar_wait_guard(&empty_a[sa_i], pha ^ 1u);
                            mbar_expect_tx(&full_a[sa_i], (uint32_t)C::A_SLOT);
                            tma_load_3d(smem + sa_i * C::A_SLOT, &map_a, kb * 128, b * a.M + I * BLK_C, p, &full_a[sa_i]);
                            if (++sa_i == NA) { sa_i = 0; pha ^= 1u; }
                        }
                    }
                } else {
                    for (int kb = lo; kb < hi; kb++) {
                        mbar_wait_guard(&empty[st], ph ^ 1u);
                        mbar_expect_tx(&full[st], (uint32_t)C64::STAGE_BYTES);
                        uint8_t* sa = smem + st * C64::STAGE_BYTES;
                        tma_load_3d(sa, &map_a, kb * BLK_K, b * a.M + I * BLK_C, 0, &full[st]);
                        tma_load_3d(sa + C64::A_BYTES, &map_b, kb * BLK_K, b * a.N + J * BLK_R, 0, &full[st]);
                        if (++st == STAGES) { st = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int st = 0, sa_i = 0;
            uint32_t ph = 0, pha = 0, acc_ph = 0;
            for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
                int b, I, J, lo, hi;
                if (!gemm_item(a, i, b, I, J)) continue;
                tile_krange(a.krange, I, J, nkb, lo, hi);
                mbar_wait_guard(tmem_empty, acc_ph ^ 1u);
                tc_fence_after();
                // digit p of the 128-row operand against digits q .. q+m-1 of the 64-row operand in ONE MMA of N = 64 m:
                // those digit tiles are contiguous in shared memory and their diagonals p+q .. p+q+m-1 are adjacent TMEM
                // column blocks, so the 128-row tile is read from shared memory once per 4 products instead of once per
                // product (at N = 64 the operand reads, 192 B/clk, exceed what shared memory delivers)
                if (K128) {
                    const int kb0 = lo >> 1, kb1 = (hi + 1) >> 1;
                    for (int kb = kb0; kb < kb1; kb++) {
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
                                const uint32_t accumulate = (kb > kb0 || ks > 0 || p > 0) ? 1u : 0u;
#pragma unroll
                                for (int q = 0; q + p < S; q += 4) {
                                    const int m = (S - p - q) < 4 ? (S - p - q) : 4;
                                    const uint64_t db = db0 + (uint64_t)((q * (BLK_R * 128) + ks * UMMA_K) >> 4);
                                    umma_i8(tmem_base + (uint32_t)((p + q) * BLK_R), da, db, idesc_i8(BLK_C, BLK_R * m), accumulate);
                                }
                            }
                            umma_commit(&empty_a[sa_i]);
                            if (++sa_i == NA) { sa_i = 0; pha ^= 1u; }
                        }
                        umma_commit(&empty[st]);
                        if (kb == kb1 - 1) umma_commit(tmem_full);
                        if (++st == STAGES) { st = 0; ph ^= 1u; }
                    }
                } else {
                    for (int kb = lo; kb < hi; kb++) {
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
                                const uint32_t accumulate = (kb > lo || ks > 0 || p > 0) ? 1u : 0u;
#pragma unroll
                                for (int q = 0; q + p < S; q += 4) {
                                    const int m = (S - p - q) < 4 ? (S - p - q) : 4;
                                    const uint64_t db = db0 + (uint64_t)((q * (BLK_R * BLK_K) + ks * UMMA_K) >> 4);
                                    umma_i8(tmem_base + (uint32_t)((p + q) * BLK_R), da, db, idesc_i8(BLK_C, BLK_R * m), accumulate);
                                }
                            }
                        }
                        umma_commit(&empty[st]);
                        if (kb == hi - 1) umma_commit(tmem_full);
                        if (++st == STAGES) { st = 0; ph ^= 1u; }
                    }
                }
                acc_ph ^= 1u;
            }
        }
    } else {
        // epilogue warps 2..9: TMEM lane quarter = warp % 4 (hardware rule), column half = (warp - 2) / 4
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t acc_ph = 0;
        constexpr double LO_W = 1.0 / (double)(1ull << (W * (S - 4)));
        // fused gradient contraction: running sums of this thread for the system it is working on
        constexpr int GD = GK >= 0 ? GRAD_D : 2;
        double gacc[GD + 2];
        int g_sys = -1;
        const bool do_grad = GK >= 0 && a.cmode == C_GRAD;
        auto grad_flush = [&]() {
            if (g_sys < 0) return;
            double* out = a.g.partial + ((int64_t)g_sys * a.g.slots + (int64_t)blockIdx.x * EPI_WARPS + (warp - 2)) * (GD + 2);
#pragma unroll
            for (int k = 0; k < GD + 2; k++) {
                const double v = warp_sum(gacc[k]);
                if (lane == 0) out[k] = v;
                gacc[k] = 0.0;
            }
        };
#pragma unroll
        for (int k = 0; k < GD + 2; k++) gacc[k] = 0.0;
        for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
            int b, I, J;
            if (!gemm_item(a, i, b, I, J)) continue;
            const int row = I * BLK_C + quarter * 32 + lane;
            // 2^(ea - 40): row scale and the weight 256^-5 of the high diagonal group
            const double row_scale = pow2d(__ldg(a.expA + (int64_t)b * a.M + row) - 2 * W * 5 / 2);
            const int* eb = a.expB + (int64_t)b * a.N + J * BLK_R + half * 32;
            double* crow = a.C + (int64_t)b * a.strideC + (int64_t)row * a.ldc + J * BLK_R + half * 32;
            mbar_wait_guard(tmem_full, acc_ph);
            tc_fence_after();
            // phase 1: this warp's 32 columns of all S diagonals -> FP64 (exact int64 recombination); TMEM is handed
            // back before any global memory is touched, so the next item's MMAs overlap the stores below
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
            // phase 2: scale and write the FP64 tile row
#pragma unroll
            for (int c = 0; c < 32; c++) z[c] = z[c] * row_scale * pow2d(__ldg(eb + c));
            if (GK >= 0 && do_grad) {
                if (b != g_sys) {                       // items are ordered by system: one flush per (CTA, system)
                    grad_flush();
                    g_sys = b;
                }
                const double* th = a.g.theta + (int64_t)b * (a.g.d + 2);
                const double g_sigma2 = __ldg(th + a.g.d), g_nugget = __ldg(th + a.g.d + 1);
                const double* al = a.g.alpha + (int64_t)b * a.g.ldv;
                // the tile's 64 columns: coordinates pre-scaled by sqrt(w_k) and alpha, staged once for the eight
                // epilogue warps (they all work on the same item); barrier 1 = the 256 epilogue threads
                asm volatile("bar.sync 1, 256;" ::: "memory");          // the previous tile's readers are done
                {
                    const int et = threadIdx.x - 64;
                    for (int e = et; e < GD * BLK_R; e += EPI_WARPS * 32) {
                        const int c = e / GD, k = e - c * GD;           // [column][k]: two coordinates per 16-byte read
                        gx_s[e] = (k < a.g.d) ? sqrt(__ldg(th + k)) * __ldg(a.g.X + (int64_t)(J * BLK_R + c) * a.g.d + k) : 0.0;
                    }
                    if (et < BLK_R) ga_s[et] = __ldg(al + J * BLK_R + et);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int gi = row;
                const bool row_ok = gi < a.g.n_valid;
                const double ai = __ldg(al + gi);
                double xi[GD];
#pragma unroll
                for (int k = 0; k < GD; k++)
                    xi[k] = (k < a.g.d) ? sqrt(__ldg(th + k)) * __ldg(a.g.X + (int64_t)gi * a.g.d + k) : 0.0;
                const int gj0 = J * BLK_R + half * 32;
                const double2* gxh = reinterpret_cast<const double2*>(gx_s + half * 32 * GD);
                const double* gah = ga_s + half * 32;
#ifdef EXQ_GRAD_SKIP_MATH
                if (false)
#endif
#pragma unroll 1
                for (int c0 = 0; c0 < 32; c0 += GRAD_ILP) {
                    // ONE copy of the pair arithmetic with GRAD_ILP pairs in flight (the FP64 chains of exp / sqrt need
                    // that many independent instructions per warp: there are only two epilogue warps per scheduler).
                    // The loop always works on z[0 .. GRAD_ILP) and rotates the array down at the end of the iteration,
                    // so z[] keeps constant indices (registers).  The coordinate differences are formed twice (for r^2
                    // and for the contraction) instead of being kept, and sqrt(w_k) is folded into the coordinates:
                    // d r^2 / d raw_k is then just (dx'_k)^2.  r2[] holds r^2, then f_ij = (K^-1 - a a^T)_ij sigma2 dcorr/dr^2.
                    double r2[GRAD_ILP];
#pragma unroll
                    for (int u = 0; u < GRAD_ILP; u++) r2[u] = 0.0;
#pragma unroll
                    for (int k = 0; k < GD; k += 2) {
#pragma unroll
                        for (int u = 0; u < GRAD_ILP; u++) {
                            const double2 xj = gxh[(c0 + u) * (GD / 2) + (k >> 1)];
                            const double d0 = xi[k] - xj.x, d1 = xi[k + 1] - xj.y;
                            r2[u] = fma(d0, d0, r2[u]);
                            r2[u] = fma(d1, d1, r2[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < GRAD_ILP; u++) {
                        const int gj = gj0 + c0 + u;
                        const double wij = z[u] - ai * gah[c0 + u];
                        const double wv = (row_ok && gj < gi) ? wij * g_sigma2 : 0.0;   // strictly lower triangle, no padding
                        if (row_ok && gj == gi) {                           // dK_ii / d ln sigma2 = sigma2, / d ln nugget = nugget
                            gacc[GD] = fma(0.5 * wij, g_sigma2, gacc[GD]);
                            gacc[GD + 1] = fma(0.5 * wij, g_nugget, gacc[GD + 1]);
                        }
                        double cv, dv;
                        if (GK == EXQ_KERNEL_MATERN52) {
                            const double s5 = sqrt_nonneg(5.0 * r2[u]), e = exp_nonpos(-s5);
                            cv = (1.0 + s5 + (5.0 / 3.0) * r2[u]) * e;
                            dv = (-5.0 / 6.0) * (1.0 + s5) * e;
                        } else {
                            cv = exp_nonpos(-0.5 * r2[u]);
                            dv = -0.5 * cv;
                        }
                        gacc[GD] = fma(wv, cv, gacc[GD]);
                        r2[u] = wv * dv;
                    }
#pragma unroll
                    for (int k = 0; k < GD; k += 2) {
#pragma unroll
                        for (int u = 0; u < GRAD_ILP; u++) {
                            const double2 xj = gxh[(c0 + u) * (GD / 2) + (k >> 1)];
                            const double d0 = xi[k] - xj.x, d1 = xi[k + 1] - xj.y;
                            gacc[k] = fma(r2[u] * d0, d0, gacc[k]);
                            gacc[k + 1] = fma(r2[u] * d1, d1, gacc[k + 1]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 32 - GRAD_ILP; c++) z[c] = z[c + GRAD_ILP];
                }
                acc_ph ^= 1u;
                continue;
            }
            double2* cp = reinterpret_cast<double2*>(crow);
            if (a.cmode == C_STORE) {
#pragma unroll
                for (int c = 0; c < 16; c++) cp[c] = make_double2(z[2 * c], z[2 * c + 1]);
                if (a.colmax) {
                    // column maxima of the tile (32 rows of this warp): one redux per column, lane c keeps column c
                    uint32_t mine = 0u;
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const uint32_t m = __reduce_max_sync(0xffffffffu, (uint32_t)__double2hiint(z[c]) & 0x7fffffffu);
                        if (lane == c) mine = m;
                    }
                    atomicMax(a.colmax + (int64_t)b * a.colmax_stride + J * BLK_R + half * 32 + lane, mine);
                }
            } else if (a.cmode == C_NEG) {
#pragma unroll
                for (int c = 0; c < 16; c++) cp[c] = make_double2(-z[2 * c], -z[2 * c + 1]);
            } else {   // C_SUB
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    double2 o = cp[c];
                    cp[c] = make_double2(o.x - z[2 * c], o.y - z[2 * c + 1]);
                }
            }
            acc_ph ^= 1u;
        }
        if (GK >= 0 && do_grad) grad_flush();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// sum the per-(CTA, epilogue warp) partials of the fused gradient epilogue in fixed order -> grad[b][0..d-1], [d], [d+1]
__global__ void grad_reduce_slots_kernel(const double* __restrict__ partial, int slots, int d, double* __restrict__ grad) {
    const int b = blockIdx.x, k = threadIdx.x;
    if (k >= d + 2) return;
    const int src = (k < d) ? k : (GRAD_D + (k - d));
    double s = 0.0;
    for (int t = 0; t < slots; t++) s += partial[((int64_t)b * slots + t) * (GRAD_D + 2) + src];
    grad[(int64_t)b * (d + 2) + k] = s;
}

template <int S>
inline cudaError_t set_gemm_attrs() {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<S, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GCfg<S>::SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_kernel<S, EXQ_KERNEL_MATERN52>, cudaFuncAttributeMaxDynamicSharedMemorySize, GCfg<S>::SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_kernel<S, EXQ_KERNEL_SQEXP>, cudaFuncAttributeMaxDynamicSharedMemorySize, GCfg<S>::SMEM_BYTES);
    return e;
}

// One operand of an int8 GEMM: exponents + digit planes for a batch of B matrices (rows x K each)
// sumsq != nullptr (transposed operands only): row exponents from the column sums of squares instead of a pass over src
// colmax != nullptr (transposed operands only): row exponents from the column maxima the producing GEMM collected
template <int S>
inline void slice_operand(const double* src, int64_t ld, int64_t stride_sys, int rows, int K, int B, int trans, int mask,
                          int* row_exp, int8_t* planes, cudaStream_t stream, const double* sumsq = nullptr,
                          int64_t sumsq_stride = 0, const uint32_t* colmax = nullptr, int64_t colmax_stride = 0) {
    const int64_t plane_bytes = (int64_t)B * rows * K;
    if (trans && colmax) {
        exponent_from_colmax_kernel<<<dim3((rows + 255) / 256, 1, B), 256, 0, stream>>>(colmax, colmax_stride, rows, row_exp);
        slice_cols_kernel<S><<<dim3(rows / 64, K / 64, B), 256, 0, stream>>>(src, ld, stride_sys, rows, K, mask, row_exp, planes,
                                                                             plane_bytes);
        return;
    }
    if (trans && sumsq) {
        exponent_from_sumsq_kernel<<<dim3((rows + 255) / 256, 1, B), 256, 0, stream>>>(sumsq, sumsq_stride, rows, row_exp);
        slice_cols_kernel<S><<<dim3(rows / 64, K / 64, B), 256, 0, stream>>>(src, ld, stride_sys, rows, K, mask, row_exp, planes,
                                                                             plane_bytes);
        return;
    }
    if (!trans) {
        const dim3 grid((rows + 7) / 8, 1, B);
        if (EXQ_OZ_SLICE_REG && K == 2048)
            slice_rows_reg_kernel<S, 16><<<grid, 256, 0, stream>>>(src, ld, stride_sys, rows, mask, row_exp, planes, plane_bytes);
        else if (EXQ_OZ_SLICE_REG && K == 1024)
            slice_rows_reg_kernel<S, 8><<<grid, 256, 0, stream>>>(src, ld, stride_sys, rows, mask, row_exp, planes, plane_bytes);
        else if (EXQ_OZ_SLICE_REG && K == 512)
            slice_rows_reg_kernel<S, 4><<<grid, 256, 0, stream>>>(src, ld, stride_sys, rows, mask, row_exp, planes, plane_bytes);
        else
            slice_rows_kernel<S><<<grid, 256, 0, stream>>>(src, ld, stride_sys, rows, K, mask, row_exp, planes, plane_bytes);
    } else {
        operand_exponent_kernel<<<dim3((rows + 31) / 32, 1, B), 256, 0, stream>>>(src, ld, stride_sys, rows, K, 1, mask, row_exp);
        slice_cols_kernel<S><<<dim3(rows / 64, K / 64, B), 256, 0, stream>>>(src, ld, stride_sys, rows, K, mask, row_exp, planes,
                                                                             plane_bytes);
    }
}

// C (op)= A B^T from digit planes: planesA [S][B*M][K], planesB [S][B*N][K]
// grad_kernel_id >= 0 with a.cmode == C_GRAD: the instantiation that carries the fused gradient epilogue
template <int S>
inline int launch_gemm(const int8_t* planesA, const int8_t* planesB, const GArgs& a, int sm_count, cudaStream_t stream,
                       int grad_kernel_id = -1) {
    CUtensorMap ma, mb;
    int rc;
    if (EXQ_OZ_K128 && (a.K % 128)) return -2;
    // K128: one digit tile of the 128-row operand per box, all S digit tiles of the 64-row operand per box
    if ((rc = make_plane_map(&ma, planesA, (int64_t)a.B * a.M, a.K, S, BLK_C, OZ_KB, EXQ_OZ_K128 ? 1 : S))) return rc;
    if ((rc = make_plane_map(&mb, planesB, (int64_t)a.B * a.N, a.K, S, BLK_R, OZ_KB, S))) return rc;
    const int items = (a.M / BLK_C) * (a.N / BLK_R) * a.B;
    const int grid = a.sched ? a.sched_grid : (items < sm_count ? items : sm_count);
    if (a.cmode == C_GRAD && grad_kernel_id == EXQ_KERNEL_MATERN52)
        gemm_kernel<S, EXQ_KERNEL_MATERN52><<<grid, THREADS, GCfg<S>::SMEM_BYTES, stream>>>(ma, mb, a);
    else if (a.cmode == C_GRAD && grad_kernel_id == EXQ_KERNEL_SQEXP)
        gemm_kernel<S, EXQ_KERNEL_SQEXP><<<grid, THREADS, GCfg<S>::SMEM_BYTES, stream>>>(ma, mb, a);
    else if (a.cmode == C_GRAD)
        return -1;
    else
        gemm_kernel<S, -1><<<grid, THREADS, GCfg<S>::SMEM_BYTES, stream>>>(ma, mb, a);
    return 0;
}

}  // namespace oz
}  // namespace exq
