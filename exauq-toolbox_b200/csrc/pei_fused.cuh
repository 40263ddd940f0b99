// Cross-covariance assembly fused with everything that used to re-read its FP64 output (VERDICT r01 "next round" 6):
// one kernel evaluates a 128-candidate x 64-training-point tile of  k(x_c, x_j) = sigma2 corr(x_c, x_j)  and
//   * writes its base-256 digit planes straight into the operand buffer of the int8 variance kernel
//     (ozaki_i8.cuh: planes[S][Mc][Np], k-contiguous, one global exponent because 0 <= k <= sigma2),
//   * accumulates the tile's contribution to the predictive mean  sum_j k_cj alpha_j  and to the repulsion product
//     prod_j (1 - k_cj / sigma2)  (exauq/core/designers.py:660-706) and stores them as per-column-tile partials.
// The FP64 chunk KcT[Mc][Np] (1 GiB per 32768 candidates at N = 4096) is never materialised: round 1 wrote it
// (8 B / entry), re-read it in the slicer (8 B) and again in the per-candidate epilogue (8 B); now 6 B of digits per
// entry cross HBM once in each direction.  fin_light_kernel then needs 2 KB of partials per candidate, not 32 KB.
// Reference arithmetic: mogp GaussianProcess.predict via exauq/core/emulators.py:649; PEICalculator.compute
// exauq/core/designers.py:522-706.
#pragma once
#include "kernels.cuh"
#include "ozaki_i8.cuh"

namespace exq {

struct AsmDigArgs {
    const double* Xc;      // candidates of this chunk [mc_pad][d]
    const double* X;       // training inputs [Np][d] (zero padded)
    const double* theta;   // [d+2]
    const double* alpha;   // [Np]
    int8_t* planes;        // [S][mc_pad][Np]
    double* PM;            // [Np/64][ldp] partial means
    double* PR;            // [Np/64][ldp] partial repulsion products
    int64_t ldp;
    int mc_pad, m_valid, N, Np, d, kernel_id;
    int cand_exp;          // global exponent of the candidate side (oz::scale_exponent(sigma2))
};

template <int KID, int S>
__global__ void __launch_bounds__(ASM_THREADS, EXQ_ASM_MIN_CTAS) assemble_digits_kernel(AsmDigArgs p) {
    const int it = blockIdx.y, jt = blockIdx.x;
    extern __shared__ __align__(16) double asm_smem[];
    const int d = p.d;
    double* Xi_s = asm_smem;                   // [128][d]
    double* Xj_s = Xi_s + 128 * d;             // [64][d]
    double* XjT = Xj_s + ASM_BN * d;           // [d][68], column c of the tile at position (c >> 2) + 16 (c & 3)
    double* w_s = XjT + d * ASM_XT_LD;         // [d+2]
    double* al_s = w_s + (d + 2);              // [64] alpha of the tile's columns, same permutation
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const double* Xi = p.Xc + (int64_t)it * 128 * d;
    const double* Xj = p.X + (int64_t)jt * ASM_BN * d;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bi = 128u * d * sizeof(double), bj = (uint32_t)ASM_BN * d * sizeof(double);
        mbar_expect_tx(&bar, bi + bj);
        tma_bulk_g2s(Xi_s, Xi, bi, &bar);
        tma_bulk_g2s(Xj_s, Xj, bj, &bar);
    }
    if (tid < d + 2) w_s[tid] = p.theta[tid];
    if (tid < ASM_BN) al_s[(tid >> 2) + 16 * (tid & 3)] = p.alpha[jt * ASM_BN + tid];
    mbar_wait(&bar, 0);
    for (int e = tid; e < ASM_BN * d; e += ASM_THREADS) {
        int pt = e / d, k = e - pt * d;
        XjT[k * ASM_XT_LD + (pt >> 2) + 16 * (pt & 3)] = Xj_s[e];
    }
    __syncthreads();

    // thread (ty, tx): candidates ty*8 .. ty*8+7, training columns 4 tx .. 4 tx + 3 (consecutive k: one packed word per plane)
    const int ty = tid >> 4, tx = tid & 15;
    constexpr bool prod = (KID == EXQ_KERNEL_PRODMAT52);
    double acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = prod ? 1.0 : 0.0;
    for (int k = 0; k < d; k++) {
        const double w = w_s[k];
        double xi[8], xj[4];
#pragma unroll
        for (int a = 0; a < 8; a++) xi[a] = Xi_s[(ty * 8 + a) * d + k];
#pragma unroll
        for (int b = 0; b < 4; b++) xj[b] = XjT[k * ASM_XT_LD + tx + 16 * b];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const double df = xi[a] - xj[b];
                if (!prod) acc[a][b] = fma(w * df, df, acc[a][b]);
                else acc[a][b] *= mat52_1d_t<EXQ_KERNEL_PRODMAT52>(w * df * df);
            }
    }
    const double sigma2 = w_s[d];
    const double inv_s2 = 1.0 / sigma2;
    const double sc = oz::pow2d(8 * S - p.cand_exp);
    constexpr unsigned long long BIAS = 0x8080808080808080ull >> (8 * (8 - S));
    double al[4];
#pragma unroll
    for (int b = 0; b < 4; b++) al[b] = al_s[tx + 16 * b];
    const int gj0 = jt * ASM_BN + 4 * tx;
    const int64_t plane = (int64_t)p.mc_pad * p.Np;
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const int gi = it * 128 + ty * 8 + a;
        uint32_t packed[S];
#pragma unroll
        for (int q = 0; q < S; q++) packed[q] = 0u;
        double sm = 0.0, rp = 1.0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            double v = 0.0;
            if (gi < p.m_valid && gj0 + b < p.N) v = sigma2 * (prod ? acc[a][b] : corr_from_r2_t<(prod ? 0 : KID)>(acc[a][b]));
            sm = fma(v, al[b], sm);
            rp *= (1.0 - div_by(v, sigma2, inv_s2));
            const unsigned long long x = (unsigned long long)__double2ll_rn(v * sc) + BIAS;
#pragma unroll
            for (int q = 0; q < S; q++) packed[q] |= (((uint32_t)(x >> (8 * (S - 1 - q))) & 0xffu) ^ 0x80u) << (8 * b);
        }
#pragma unroll
        for (int q = 0; q < S; q++)
            *reinterpret_cast<uint32_t*>(p.planes + q * plane + (int64_t)gi * p.Np + gj0) = packed[q];
        // the 16 threads tx = 0..15 of a row are 16 consecutive lanes: fixed-order butterfly
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            sm += __shfl_xor_sync(0xffffffffu, sm, o);
            rp *= __shfl_xor_sync(0xffffffffu, rp, o);
        }
        if (tx == 0) {
            p.PM[(int64_t)jt * p.ldp + gi] = sm;
            p.PR[(int64_t)jt * p.ldp + gi] = rp;
        }
    }
}

struct FinLightArgs {
    const double *PM, *PR;   // [nM][ldp]
    const double* P;         // [nP][ldp] partial column sums of squares of the variance product
    int64_t ldp;
    int nM, nP;
    const double* theta;
    const double* Xc;        // [Mc][d]
    const double* Rp;        // [n_rep][d]
    int rp_in_smem;          // the repulsion points fit the dynamic shared memory of the launch
    int n_rep, d, kernel_id, m_valid;
    double tmax;
    int want_pei;
    double *mean, *var, *pei;
};

// FIN_LANES threads per candidate (a candidate's lanes are neighbours in a warp): each lane sums / multiplies every
// FIN_LANES-th per-tile partial and every FIN_LANES-th of the other repulsion points (pseudopoints and earlier design
// points), a fixed-order butterfly combines them, lane 0 finishes expected improvement and PEI.  (One thread per
// candidate left 89 % of the warp slots empty at 32768 candidates per chunk: 206 us per chunk, latency-bound.)
constexpr int FIN_LANES = 4;
__global__ void __launch_bounds__(128) fin_light_kernel(FinLightArgs p) {
    extern __shared__ double fl_smem[];   // w[d+2], then Rp[n_rep][d] when it fits
    const int d = p.d;
    double* w_s = fl_smem;
    double* rp_s = fl_smem + (d + 2);
    for (int e = threadIdx.x; e < d + 2; e += blockDim.x) w_s[e] = p.theta[e];
    if (p.rp_in_smem)
        for (int e = threadIdx.x; e < p.n_rep * d; e += blockDim.x) rp_s[e] = p.Rp[e];
    __syncthreads();
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = gt / FIN_LANES, sub = gt % FIN_LANES;
    const bool live = c < p.m_valid;
    const int cc = live ? c : 0;                      // idle lanes of the last warp shadow candidate 0 (no divergent shuffles)
    double sm = 0.0, rep = 1.0, q = 0.0;
    for (int r = sub; r < p.nM; r += FIN_LANES) {
        sm += p.PM[(int64_t)r * p.ldp + cc];
        rep *= p.PR[(int64_t)r * p.ldp + cc];
    }
    for (int r = sub; r < p.nP; r += FIN_LANES) q += p.P[(int64_t)r * p.ldp + cc];
    if (p.want_pei) {
        const double* xc = p.Xc + (int64_t)cc * d;
        const double* rp = p.rp_in_smem ? rp_s : p.Rp;
        for (int r = sub; r < p.n_rep; r += FIN_LANES) rep *= (1.0 - point_corr(p.kernel_id, xc, rp + (int64_t)r * d, w_s, d));
    }
#pragma unroll
    for (int o = 1; o < FIN_LANES; o <<= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
        rep *= __shfl_xor_sync(0xffffffffu, rep, o);
    }
    if (!live || sub != 0) return;
    const double sigma2 = w_s[d], nugget = w_s[d + 1];
    const double var = fmax(sigma2 + nugget - q, 0.0);
    if (p.mean) p.mean[c] = sm;
    if (p.var) p.var[c] = var;
    if (p.want_pei) p.pei[c] = expected_improvement(sm, var, p.tmax) * rep;
}

}  // namespace exq
