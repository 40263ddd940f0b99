// Covariance assembly, O(N^2) vector kernels, LOO / PEI / argmax epilogues and the
// log-posterior gradient contraction.  Reference arithmetic cited per kernel.
#pragma once
#include "exq_common.cuh"
#include <math_constants.h>

namespace exq {

// theta layout per system (device, doubles): w[0..d-1] = exp(corr_raw) = 1/l^2, then
// sigma2 (process variance), nugget (value added to the diagonal).
#define EXQ_THETA_LEN(d) ((d) + 2)

// =====================================================================================
// K1  covariance assembly.
//   SYM  : K[i][j] = sigma2 * corr(x_i, x_j) (+ nugget on the diagonal), lower tiles only,
//          identity on the padding  (mogp GaussianProcess.get_K_matrix + nugget;
//          call site exauq/core/emulators.py:447-448)
//   CROSS: KcT[c][j] = sigma2 * corr(xc_c, x_j), zero on the padding of j
//          (AbstractGaussianProcess.covariance_matrix, exauq/core/modelling.py:1042-1080)
// Point tiles (128 points x d, contiguous) are staged with one TMA bulk copy each.
// =====================================================================================
constexpr int GRAD_XT_LD = 132;
#ifndef EXQ_ASM_MIN_CTAS
#define EXQ_ASM_MIN_CTAS 2
#endif
constexpr int ASM_THREADS = 256;
constexpr int ASM_BN = 64;          // CTA tile: 128 rows x 64 columns, 8 x 4 entries per thread
constexpr int ASM_XT_LD = ASM_BN + 4;

struct AsmArgs {
    const double* Xi;      // row points  [rows_pad][d]
    const double* Xj;      // col points  [Np][d]
    const double* theta;   // [B][d+2]
    double* out;           // [rows_pad][ld]
    int64_t ld, strideOut, strideXi, strideXj, strideTheta;
    int n_i, n_j;          // valid rows / cols (rest is padding)
    int d, kernel_id, sym;
    int scale_sigma2;      // 0: plain correlation (no sigma2, no nugget)
};

// KID >= 0: kernel fixed at compile time + branch-free exp / sqrt (exq_common.cuh); KID = -1: the generic
// version with the library functions and a run-time kernel id (EXQ_FAST_CORR=0, kept as the A/B reference)
template <int KID>
__global__ void __launch_bounds__(ASM_THREADS, EXQ_ASM_MIN_CTAS) assemble_kernel(AsmArgs p) {
    const int it = blockIdx.y, jt = blockIdx.x;
    if (p.sym && jt * ASM_BN > it * 128 + 127) return;
    extern __shared__ __align__(16) double asm_smem[];
    const int d = p.d;
    double* Xi_s = asm_smem;                   // [128][d]
    double* Xj_s = Xi_s + 128 * d;             // [64][d]
    double* XjT = Xj_s + ASM_BN * d;           // [d][68]
    double* w_s = XjT + d * ASM_XT_LD;         // [d+2]
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const double* Xi = p.Xi + (int64_t)blockIdx.z * p.strideXi + (int64_t)it * 128 * d;
    const double* Xj = p.Xj + (int64_t)blockIdx.z * p.strideXj + (int64_t)jt * ASM_BN * d;
    const double* theta = p.theta + (int64_t)blockIdx.z * p.strideTheta;
    double* out = p.out + (int64_t)blockIdx.z * p.strideOut;

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
    if (tid < d + 2) w_s[tid] = theta[tid];
    mbar_wait(&bar, 0);
    for (int e = tid; e < ASM_BN * d; e += ASM_THREADS) {
        int pt = e / d, k = e - pt * d;
        XjT[k * ASM_XT_LD + pt] = Xj_s[e];
    }
    __syncthreads();

    const int ty = tid >> 4, tx = tid & 15;
    double acc[8][4];
    const bool prod = KID >= 0 ? (KID == EXQ_KERNEL_PRODMAT52) : (p.kernel_id == EXQ_KERNEL_PRODMAT52);
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
        if (!prod) {
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    double df = xi[a] - xj[b];
                    acc[a][b] = fma(w * df, df, acc[a][b]);
                }
        } else {
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    double df = xi[a] - xj[b];
                    acc[a][b] *= KID >= 0 ? mat52_1d_t<EXQ_KERNEL_PRODMAT52>(w * df * df) : mat52_1d(w * df * df);
                }
        }
    }
    const double sigma2 = p.scale_sigma2 ? w_s[d] : 1.0;
    const double nugget = p.scale_sigma2 ? w_s[d + 1] : 0.0;
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const int gi = it * 128 + ty * 8 + a;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int gj = jt * ASM_BN + tx + 16 * b;
            double v;
            if (gi < p.n_i && gj < p.n_j) {
                double cv;
                if (KID >= 0) cv = prod ? acc[a][b] : corr_from_r2_t<(KID >= 0 ? KID : 0)>(acc[a][b]);
                else cv = prod ? acc[a][b] : corr_from_r2(p.kernel_id, acc[a][b]);
                v = sigma2 * cv;
                if (p.sym && gi == gj) v = sigma2 + nugget;
            } else {
                v = (p.sym && gi == gj) ? 1.0 : 0.0;
            }
            out[(int64_t)gi * p.ld + gj] = v;
        }
    }
}

// =====================================================================================
// O(N^2) triangular mat-vecs with the inverse factor.
//   t = Linv y ;  alpha = Linv^T t = K^-1 y ;  dinv_j = sum_i Linv[i][j]^2 = [K^-1]_jj
// alpha is mogp's Kinv_t (GaussianProcess.fit); dinv feeds the closed-form LOO sweep.
// =====================================================================================
__global__ void __launch_bounds__(256) trmv_lower_kernel(const double* __restrict__ Linv, int64_t ld,
                                                         int64_t strideL, const double* __restrict__ y,
                                                         int64_t strideY, double* __restrict__ tvec,
                                                         int64_t strideT, int Np) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= Np) return;
    const double* row = Linv + (int64_t)blockIdx.z * strideL + (int64_t)i * ld;
    const double* yy = y + (int64_t)blockIdx.z * strideY;
    double s = 0.0;
    for (int k = lane; k <= i; k += 32) s += row[k] * yy[k];
    s = warp_sum(s);
    if (lane == 0) tvec[(int64_t)blockIdx.z * strideT + i] = s;
}

// Row-split version: blockIdx.y handles rows [y*chunk, (y+1)*chunk) and writes partial sums
// part[b][y][0][j] (alpha) and part[b][y][1][j] (dinv); trmv_t_reduce_kernel adds the splits in
// fixed order.  (One block per 32 columns alone leaves most SMs idle for a single system.)
constexpr int TRMV_T_SPLITS = 8;
__global__ void __launch_bounds__(256) trmv_lower_t_kernel(const double* __restrict__ Linv, int64_t ld,
                                                           int64_t strideL, const double* __restrict__ tvec,
                                                           int64_t strideT, double* __restrict__ part, int Np) {
    __shared__ double red_a[8][33], red_d[8][33];
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const int j = blockIdx.x * 32 + tx;
    const double* L = Linv + (int64_t)blockIdx.z * strideL;
    const double* tv = tvec + (int64_t)blockIdx.z * strideT;
    const int chunk = (Np + TRMV_T_SPLITS - 1) / TRMV_T_SPLITS;
    const int r_lo = max((int)blockIdx.y * chunk, (int)blockIdx.x * 32), r_hi = min(Np, ((int)blockIdx.y + 1) * chunk);
    double sa = 0.0, sd = 0.0;
#pragma unroll 4
    for (int i = r_lo + ty; i < r_hi; i += 8) {
        double v = (i >= j) ? L[(int64_t)i * ld + j] : 0.0;
        sa = fma(v, tv[i], sa);
        sd = fma(v, v, sd);
    }
    red_a[ty][tx] = sa;
    red_d[ty][tx] = sd;
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int k = 1; k < 8; k++) { sa += red_a[k][tx]; sd += red_d[k][tx]; }
        double* out = part + ((int64_t)blockIdx.z * TRMV_T_SPLITS + blockIdx.y) * 2 * Np;
        out[j] = sa;
        out[Np + j] = sd;
    }
}

__global__ void trmv_t_reduce_kernel(const double* __restrict__ part, double* __restrict__ alpha,
                                     double* __restrict__ dinv, int64_t strideV, int Np) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Np) return;
    const double* p = part + (int64_t)blockIdx.z * TRMV_T_SPLITS * 2 * Np;
    double sa = 0.0, sd = 0.0;
#pragma unroll
    for (int y = 0; y < TRMV_T_SPLITS; y++) { sa += p[(int64_t)y * 2 * Np + j]; sd += p[(int64_t)y * 2 * Np + Np + j]; }
    alpha[(int64_t)blockIdx.z * strideV + j] = sa;
    dinv[(int64_t)blockIdx.z * strideV + j] = sd;
}

// logdet = 2 sum_{i<N} log L_ii ; quad = y^T alpha.   out[b] = {logdet, quad};  dstat[b] = {min_i L_ii, max_i L_ii}
// (max/min)^2 is a lower bound of cond_2(K): the conditioning proxy of the int8 digit guard (exq_api.cu)
__global__ void __launch_bounds__(256) logdet_quad_kernel(const double* __restrict__ A, int64_t ld,
                                                          int64_t strideA, const double* __restrict__ y,
                                                          int64_t strideY, const double* __restrict__ alpha,
                                                          int64_t strideV, int N, double* __restrict__ out,
                                                          double* __restrict__ dstat) {
    __shared__ double r1[8], r2[8], r3[8], r4[8];
    const double* a = A + (int64_t)blockIdx.x * strideA;
    const double* yy = y + (int64_t)blockIdx.x * strideY;
    const double* al = alpha + (int64_t)blockIdx.x * strideV;
    double s1 = 0.0, s2 = 0.0, lo = CUDART_INF, hi = 0.0;
    for (int i = threadIdx.x; i < N; i += 256) {
        const double l = a[(int64_t)i * ld + i];
        s1 += log(l);
        s2 += yy[i] * al[i];
        lo = fmin(lo, l);
        hi = fmax(hi, l);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; r3[threadIdx.x >> 5] = lo; r4[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) { s1 += r1[k]; s2 += r2[k]; lo = fmin(lo, r3[k]); hi = fmax(hi, r4[k]); }
        out[2 * blockIdx.x] = 2.0 * s1;
        out[2 * blockIdx.x + 1] = s2;
        dstat[2 * blockIdx.x] = lo;
        dstat[2 * blockIdx.x + 1] = hi;
    }
}

// y = M x for a dense row-major n x n matrix (one warp per row): power iteration of the kinv condition estimate
__global__ void __launch_bounds__(256) gemv_kernel(const double* __restrict__ Mx, int64_t ld, const double* __restrict__ x,
                                                   double* __restrict__ yout, int n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= n) return;
    const double* row = Mx + (int64_t)i * ld;
    double s = 0.0;
    for (int k = lane; k < n; k += 32) s = fma(row[k], x[k], s);
    s = warp_sum(s);
    if (lane == 0) yout[i] = s;
}

// x <- x / |x|_2 ; norm_out[0] = |x|_2 (one block)
__global__ void __launch_bounds__(256) normalize_kernel(double* __restrict__ x, int n, double* __restrict__ norm_out) {
    __shared__ double red[8];
    __shared__ double nrm;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s = fma(x[i], x[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) s += red[k];
        nrm = sqrt(s);
        norm_out[0] = nrm;
    }
    __syncthreads();
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (int i = threadIdx.x; i < n; i += 256) x[i] *= inv;
}

// =====================================================================================
// K3  closed-form leave-one-out sweep:
//   refit on data \ {i} with the parent's hyperparameters and predict at x_i
//   (exauq/core/designers.py:263-272, 295-370) ==  mu_i = y_i - alpha_i / [K^-1]_ii,
//   var_i = 1 / [K^-1]_ii (clipped at 0 as mogp predict does), and the NES error of
//   exauq/core/modelling.py:754-825 against y_i.
// =====================================================================================
__device__ __forceinline__ double nes_error(double mean, double var, double observed) {
    double sq = (mean - observed) * (mean - observed);
    double num = var + sq;
    double den = sqrt(2.0 * (var * var) + 4.0 * var * sq);
    if (den == 0.0) return (num == 0.0) ? 0.0 : CUDART_INF;
    return num / den;
}

__global__ void loo_kernel(const double* __restrict__ y, const double* __restrict__ alpha,
                           const double* __restrict__ dinv, int N, double* __restrict__ mean,
                           double* __restrict__ var, double* __restrict__ nes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double v = 1.0 / dinv[i];
    double m = y[i] - alpha[i] * v;
    v = fmax(v, 0.0);
    mean[i] = m;
    var[i] = v;
    nes[i] = nes_error(m, v, y[i]);
}

// =====================================================================================
// K4  per-candidate epilogue after the variance GEMM.  One warp per candidate.
//   mean = k^T alpha ; var = max(sigma2 + nugget - |Linv k|^2, 0)       (mogp predict)
//   EI (exauq/core/designers.py:610-658), repulsion (:660-706), PEI = EI * repulsion (:522)
// =====================================================================================
struct FinArgs {
    const double* KcT;     // [Mc][ld]  sigma2 * corr(candidate, train)
    int64_t ld;
    const double* P;       // [nrb][ldp] partial column sums of squares
    int64_t ldp;
    int nrb;
    const double* alpha;   // [Np]
    const double* theta;   // [d+2]
    const double* Xc;      // [Mc][d] candidates of this chunk
    const double* Rp;      // [n_rep][d] other repulsion points (pseudopoints etc.)
    int n_rep;
    int N, Np, d, kernel_id;
    int m_valid;           // valid candidates in this chunk
    double tmax;           // max training target of the errors GP
    int want_pei;
    double* mean;          // [Mc] or null
    double* var;           // [Mc] or null
    double* pei;           // [Mc] or null
};

__device__ __forceinline__ double point_corr(int kernel_id, const double* __restrict__ a,
                                             const double* __restrict__ b, const double* __restrict__ w,
                                             int d) {
    if (kernel_id == EXQ_KERNEL_PRODMAT52) {
        double pr = 1.0;
        for (int k = 0; k < d; k++) { double df = a[k] - b[k]; pr *= mat52_1d(w[k] * df * df); }
        return pr;
    }
    double r2 = 0.0;
    for (int k = 0; k < d; k++) { double df = a[k] - b[k]; r2 = fma(w[k] * df, df, r2); }
    return corr_from_r2(kernel_id, r2);
}

// =====================================================================================
// Iterative refinement of alpha = K^-1 y (exq_api.cu: refine_alpha).  r = y - (sigma2 C + nugget I) alpha with the
// covariance entries recomputed from the points in FP64 (never read from the factorised copy), one 32-row slab per
// CTA, 8 column lanes per row.  With the inverse factor built from int8 digit products the first alpha carries
// errors that are small in norm but unstructured, and the predictive mean k^T alpha -- a sum with cond(K)-fold
// cancellation -- showed them (tools/guard_survey.py: 4e-8 against 7e-10 on the FP64 DMMA path at cond 4e7); one
// refinement step restores the backward-stable level.  Reference: mogp computes alpha by cho_solve in
// GaussianProcess.fit (exauq/core/emulators.py:447-448).
// =====================================================================================
__global__ void __launch_bounds__(256) kresidual_kernel(const double* __restrict__ Xall, int64_t strideX,
                                                        const double* __restrict__ thetaall, int64_t strideTheta,
                                                        const double* __restrict__ yall, int64_t strideY,
                                                        const double* __restrict__ alphaall, int64_t strideV,
                                                        double* __restrict__ rall, int N, int Np, int d, int kernel_id) {
    extern __shared__ double kr_smem[];   // Xi[32][d], w[d+2]
    double* Xi_s = kr_smem;
    double* w_s = kr_smem + 32 * d;
    const double* X = Xall + (int64_t)blockIdx.z * strideX;
    const double* theta = thetaall + (int64_t)blockIdx.z * strideTheta;
    const double* y = yall + (int64_t)blockIdx.z * strideY;
    const double* alpha = alphaall + (int64_t)blockIdx.z * strideV;
    double* r = rall + (int64_t)blockIdx.z * strideV;
    const int row0 = blockIdx.x * 32;
    for (int e = threadIdx.x; e < 32 * d; e += 256) Xi_s[e] = (row0 + e / d < N) ? X[(int64_t)row0 * d + e] : 0.0;
    if (threadIdx.x < d + 2) w_s[threadIdx.x] = theta[threadIdx.x];
    __syncthreads();
    const int lr = threadIdx.x >> 3, cl = threadIdx.x & 7;
    const int i = row0 + lr;
    const double sigma2 = w_s[d], nugget = w_s[d + 1];
    double acc = 0.0;
    if (i < N) {
        const double* xi = Xi_s + lr * d;
        for (int j = cl; j < N; j += 8) {
            double cv = (j == i) ? 1.0 : point_corr(kernel_id, xi, X + (int64_t)j * d, w_s, d);
            double kij = sigma2 * cv + ((j == i) ? nugget : 0.0);
            acc = fma(kij, alpha[j], acc);
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (cl == 0 && i < Np) r[i] = (i < N) ? y[i] - acc : 0.0;
}

// alpha += delta (padding untouched)
__global__ void axpy_kernel(double* __restrict__ alpha, const double* __restrict__ delta, int64_t strideV, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) alpha[(int64_t)blockIdx.z * strideV + i] += delta[(int64_t)blockIdx.z * strideV + i];
}

__device__ __forceinline__ double expected_improvement(double mean, double var, double tmax) {
    double sd = sqrt(var);
    if (fabs(sd) <= 1e-9) return 0.0;          // equal_within_tolerance(sd, 0), numerics.py:31-73
    double u = (mean - tmax) / sd;
    double cdf = normcdf(u);
    double pdf = exp(-0.5 * u * u) * 0.3989422804014327;   // 1/sqrt(2 pi)
    return (mean - tmax) * cdf + sd * pdf;
}

// a / b for a warp-uniform divisor with y = 1/b precomputed: one Newton correction of a*y gives
// the correctly rounded quotient (Markstein) in 3 FP64 ops instead of the ~20 of a full division.
__device__ __forceinline__ double div_by(double a, double b, double y) {
    double q = a * y;
    double r = fma(-q, b, a);
    return fma(r, y, q);
}

#ifndef EXQ_FIN_MIN_BLOCKS
#define EXQ_FIN_MIN_BLOCKS 3
#endif
#ifndef EXQ_FIN_UNROLL
#define EXQ_FIN_UNROLL 8
#endif
constexpr int FIN_UNROLL = EXQ_FIN_UNROLL;
__global__ void __launch_bounds__(256, EXQ_FIN_MIN_BLOCKS) predict_finalize_kernel(FinArgs p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + warp;
    if (c >= p.m_valid) return;
    const double* row = p.KcT + (int64_t)c * p.ld;
    const double sigma2 = p.theta[p.d], nugget = p.theta[p.d + 1];
    const double inv_s2 = 1.0 / sigma2;
    // streaming pass over the candidate's covariance row: 16-byte loads, 4 independent chains
    double sm0 = 0.0, sm1 = 0.0, rp0 = 1.0, rp1 = 1.0;
    const int n2 = p.N >> 1;
    const double2* row2 = reinterpret_cast<const double2*>(row);
    const double2* al2 = reinterpret_cast<const double2*>(p.alpha);
#pragma unroll FIN_UNROLL
    for (int j = lane; j < n2; j += 32) {
        double2 kv = __ldcs(row2 + j);
        double2 al = al2[j];
        sm0 = fma(kv.x, al.x, sm0);
        sm1 = fma(kv.y, al.y, sm1);
        rp0 *= (1.0 - div_by(kv.x, sigma2, inv_s2));
        rp1 *= (1.0 - div_by(kv.y, sigma2, inv_s2));
    }
    if ((p.N & 1) && lane == 0) {
        double kv = row[p.N - 1];
        sm0 = fma(kv, p.alpha[p.N - 1], sm0);
        rp0 *= (1.0 - div_by(kv, sigma2, inv_s2));
    }
    double sm = sm0 + sm1, rep = rp0 * rp1;
    double q = 0.0;
    for (int r = lane; r < p.nrb; r += 32) q += p.P[(int64_t)r * p.ldp + c];
    sm = warp_sum(sm);
    q = warp_sum(q);
    double var = fmax(sigma2 + nugget - q, 0.0);
    if (p.want_pei) {
        const double* xc = p.Xc + (int64_t)c * p.d;
        for (int r = lane; r < p.n_rep; r += 32)
            rep *= (1.0 - point_corr(p.kernel_id, xc, p.Rp + (int64_t)r * p.d, p.theta, p.d));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rep *= __shfl_xor_sync(0xffffffffu, rep, o);
    }
    if (lane == 0) {
        if (p.mean) p.mean[c] = sm;
        if (p.var) p.var[c] = var;
        if (p.want_pei) p.pei[c] = expected_improvement(sm, var, p.tmax) * rep;
    }
}

// sum of squares of each of the M vectors z[c][0..n) -> out[c]  (small-M variance path)
__global__ void __launch_bounds__(256) sumsq_rows_kernel(const double* __restrict__ z, int64_t ld, int n,
                                                         double* __restrict__ out) {
    __shared__ double red[8];
    const double* row = z + (int64_t)blockIdx.x * ld;
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += 256) s = fma(row[j], row[j], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) s += red[k];
        out[blockIdx.x] = s;
    }
}

// ---- argmax over M values, lowest index wins ties, NaN ignored ------------------------
struct ArgPair { double v; long long i; };
__device__ __forceinline__ bool arg_better(double v, long long i, double bv, long long bi) {
    return (v > bv) || (v == bv && i < bi);
}
__device__ __forceinline__ void arg_warp_reduce(double& v, long long& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        long long oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (arg_better(ov, oi, v, i)) { v = ov; i = oi; }
    }
}
__device__ __forceinline__ void arg_block_reduce(double& v, long long& i) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    arg_warp_reduce(v, i);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sv[warp] = v; si[warp] = i; }
    __syncthreads();
    if (warp == 0) {
        int nw = blockDim.x >> 5;
        v = (lane < nw) ? sv[lane] : -CUDART_INF;
        i = (lane < nw) ? si[lane] : 0x7fffffffffffffffLL;
        arg_warp_reduce(v, i);
    }
}

// PEI update for one newly chosen design point (exauq/core/designers.py:835: the chosen
// point joins the repulsion set; the GP is not refit) fused with the block-level argmax:
//   pei[c] *= 1 - corr(xc_c, x_new)   (skipped when xnew == null)
__global__ void __launch_bounds__(256) pei_update_argmax_kernel(double* __restrict__ pei,
                                                                const double* __restrict__ Xc, long long M,
                                                                int d, int kernel_id,
                                                                const double* __restrict__ theta,
                                                                const double* __restrict__ xnew,
                                                                ArgPair* __restrict__ partial) {
    extern __shared__ double upd_smem[];   // x_new[d], w[d]
    const bool upd = (xnew != nullptr);
    if (upd) {
        for (int k = threadIdx.x; k < d; k += blockDim.x) {
            upd_smem[k] = xnew[k];
            upd_smem[d + k] = theta[k];
        }
        __syncthreads();
    }
    double bv = -CUDART_INF;
    long long bi = 0x7fffffffffffffffLL;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < M;
         c += (long long)gridDim.x * blockDim.x) {
        double v = pei[c];
        if (upd) {
            double cr;
            if ((d & 1) == 0 && kernel_id != EXQ_KERNEL_PRODMAT52) {
                // 16-byte streaming loads of the candidate's coordinates (rows are 16-byte aligned when d is even)
                const double2* x2 = reinterpret_cast<const double2*>(Xc + c * d);
                double r2 = 0.0;
                for (int k = 0; k < d; k += 2) {
                    double2 xv = __ldcs(x2 + (k >> 1));
                    double d0 = xv.x - upd_smem[k], d1 = xv.y - upd_smem[k + 1];
                    r2 = fma(upd_smem[d + k] * d0, d0, r2);
                    r2 = fma(upd_smem[d + k + 1] * d1, d1, r2);
                }
                cr = corr_from_r2(kernel_id, r2);
            } else {
                cr = point_corr(kernel_id, Xc + c * d, upd_smem, upd_smem + d, d);
            }
            v *= (1.0 - cr);
            pei[c] = v;
        }
        if (arg_better(v, c, bv, bi)) { bv = v; bi = c; }
    }
    arg_block_reduce(bv, bi);
    if (threadIdx.x == 0) { partial[blockIdx.x].v = bv; partial[blockIdx.x].i = bi; }
}

__global__ void __launch_bounds__(256) argmax_final_kernel(const ArgPair* __restrict__ partial, int n,
                                                           ArgPair* __restrict__ out) {
    double bv = -CUDART_INF;
    long long bi = 0x7fffffffffffffffLL;
    for (int k = threadIdx.x; k < n; k += blockDim.x)
        if (arg_better(partial[k].v, partial[k].i, bv, bi)) { bv = partial[k].v; bi = partial[k].i; }
    arg_block_reduce(bv, bi);
    if (threadIdx.x == 0) { out->v = bv; out->i = bi; }
}

// =====================================================================================
// K5  gradient of the negative log-posterior w.r.t. the raw parameters (data part):
//   dL/dp = 0.5 * sum_ij (K^-1 - alpha alpha^T)_ij  dK_ij/dp
// (mogp GaussianProcess.logpost_deriv; call site exauq/utilities/mogp_fitting.py:318-324.)
//   p = corr_raw_k : dK_ij = sigma2 * corr'(r2) * w_k (x_ik - x_jk)^2
//   p = ln sigma2  : dK_ij = sigma2 * corr(r2)           (no nugget)
//   p = ln nugget  : dK_ij = nugget * delta_ij           (only reported; used when fitted)
// One CTA per lower 128 x 128 tile pair; per-CTA partial sums are reduced in fixed order.
// =====================================================================================
template <int DMAX>
__global__ void __launch_bounds__(256) grad_kernel(const double* __restrict__ Xall, int64_t strideX,
                                                   const double* __restrict__ thetaall, int64_t strideTheta,
                                                   const double* __restrict__ Kinvall, int64_t ld,
                                                   int64_t strideK, const double* __restrict__ alphaall,
                                                   int64_t strideV, int N, int d, int kernel_id,
                                                   double* __restrict__ partial /*[B][ntiles][DMAX+2]*/,
                                                   int ntiles_side) {
    const int it = blockIdx.y, jt = blockIdx.x;
    if (jt > it) return;
    extern __shared__ __align__(16) double g_smem[];
    double* Xi_s = g_smem;                  // [128][d]
    double* XjT = Xi_s + 128 * d;           // [d][132]
    double* w_s = XjT + d * GRAD_XT_LD;      // [d+2]
    double* ai_s = w_s + (d + 2);           // [128]
    double* aj_s = ai_s + 128;              // [128]
    double* red = aj_s + 128;               // [8][DMAX+2]
    const int tid = threadIdx.x;
    const double* X = Xall + (int64_t)blockIdx.z * strideX;
    const double* theta = thetaall + (int64_t)blockIdx.z * strideTheta;
    const double* Kinv = Kinvall + (int64_t)blockIdx.z * strideK;
    const double* alpha = alphaall + (int64_t)blockIdx.z * strideV;

    for (int e = tid; e < 128 * d; e += 256) {
        int pt = e / d, k = e - pt * d;
        Xi_s[e] = X[(int64_t)(it * 128) * d + e];
        XjT[k * GRAD_XT_LD + pt] = X[(int64_t)(jt * 128) * d + e];
    }
    if (tid < d + 2) w_s[tid] = theta[tid];
    if (tid < 128) { ai_s[tid] = alpha[it * 128 + tid]; aj_s[tid] = alpha[jt * 128 + tid]; }
    __syncthreads();

    const int ty = tid >> 4, tx = tid & 15;
    const double sigma2 = w_s[d], nugget = w_s[d + 1];
    double acc[DMAX + 2];
#pragma unroll
    for (int k = 0; k < DMAX + 2; k++) acc[k] = 0.0;

    for (int a = 0; a < 8; a++) {
        const int li = ty * 8 + a, gi = it * 128 + li;
        if (gi >= N) continue;
        for (int b = 0; b < 8; b++) {
            const int lj = tx + 16 * b, gj = jt * 128 + lj;
            if (gj > gi || gj >= N) continue;
            const double W = Kinv[(int64_t)gi * ld + gj] - ai_s[li] * aj_s[lj];
            if (gi == gj) {
                acc[DMAX] += 0.5 * W * sigma2;
                acc[DMAX + 1] += 0.5 * W * nugget;
                continue;
            }
            double gk[DMAX];
            double r2 = 0.0;
#pragma unroll
            for (int k = 0; k < DMAX; k++) {
                if (k < d) {
                    double df = Xi_s[li * d + k] - XjT[k * GRAD_XT_LD + lj];
                    gk[k] = w_s[k] * df * df;
                    r2 += gk[k];
                } else gk[k] = 0.0;
            }
            if (kernel_id == EXQ_KERNEL_PRODMAT52) {
                // K = prod_k m(g_k):  dK/draw_k = K * m'(g_k)/m(g_k) * g_k  (mogp ProductMat52.kernel_deriv)
                double cvp = 1.0;
#pragma unroll
                for (int k = 0; k < DMAX; k++) {
                    if (k < d) {
                        double s5 = sqrt_nonneg(5.0 * gk[k]), e = exp_nonpos(-s5);
                        double mk = (1.0 + s5 + (5.0 / 3.0) * gk[k]) * e;
                        double dk = (-5.0 / 6.0) * (1.0 + s5) * e;
                        cvp *= mk;
                        gk[k] = gk[k] * dk / mk;      // logarithmic derivative factor
                    }
                }
                const double fp = W * sigma2 * cvp;
#pragma unroll
                for (int k = 0; k < DMAX; k++) acc[k] = fma(fp, gk[k], acc[k]);
                acc[DMAX] += fp;
                continue;
            }
            double cv, dv;
            if (kernel_id == EXQ_KERNEL_MATERN52) {
                double s5 = sqrt_nonneg(5.0 * r2), e = exp_nonpos(-s5);
                cv = (1.0 + s5 + (5.0 / 3.0) * r2) * e;
                dv = (-5.0 / 6.0) * (1.0 + s5) * e;
            } else {
                cv = exp_nonpos(-0.5 * r2);
                dv = -0.5 * cv;
            }
            const double f = W * sigma2 * dv;
#pragma unroll
            for (int k = 0; k < DMAX; k++) acc[k] = fma(f, gk[k], acc[k]);
            acc[DMAX] = fma(W * sigma2, cv, acc[DMAX]);
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int k = 0; k < DMAX + 2; k++) {
        double v = warp_sum(acc[k]);
        if (lane == 0) red[warp * (DMAX + 2) + k] = v;
    }
    __syncthreads();
    if (tid < DMAX + 2) {
        double s = 0.0;
        for (int wv = 0; wv < 8; wv++) s += red[wv * (DMAX + 2) + tid];
        int tile = it * ntiles_side + jt;
        partial[((int64_t)blockIdx.z * ntiles_side * ntiles_side + tile) * (DMAX + 2) + tid] = s;
    }
}

// Same contraction for the two stationary-distance kernels (KID = Matern-5/2 or squared exponential) with four
// independent (i, j) pairs in flight per thread, no data-dependent branches in the pair loop (pairs on or above the
// diagonal, or in the padding, enter with weight 0) and the branch-free exp / sqrt of exq_common.cuh: the generic kernel
// above walks its 64 pairs one at a time through a ~60-operation dependent chain and reached 17 % of the FP64 pipe
// (ncu launch list, round 2: 1.47 ms for 15 systems at N = 4096).  Same partial-sum layout, same reduction kernel.
template <int KID, int DMAX>
__global__ void __launch_bounds__(256) grad_fast_kernel(const double* __restrict__ Xall, int64_t strideX,
                                                        const double* __restrict__ thetaall, int64_t strideTheta,
                                                        const double* __restrict__ Kinvall, int64_t ld, int64_t strideK,
                                                        const double* __restrict__ alphaall, int64_t strideV, int N, int d,
                                                        double* __restrict__ partial /*[B][ntiles][DMAX+2]*/,
                                                        int ntiles_side) {
    const int it = blockIdx.y, jt = blockIdx.x;
    if (jt > it) return;
    extern __shared__ __align__(16) double g_smem[];
    double* Xi_s = g_smem;                  // [128][d]
    double* XjT = Xi_s + 128 * d;           // [d][132]
    double* w_s = XjT + d * GRAD_XT_LD;      // [d+2]
    double* ai_s = w_s + (d + 2);           // [128]
    double* aj_s = ai_s + 128;              // [128]
    double* red = aj_s + 128;               // [8][DMAX+2]
    const int tid = threadIdx.x;
    const double* X = Xall + (int64_t)blockIdx.z * strideX;
    const double* theta = thetaall + (int64_t)blockIdx.z * strideTheta;
    const double* Kinv = Kinvall + (int64_t)blockIdx.z * strideK;
    const double* alpha = alphaall + (int64_t)blockIdx.z * strideV;

    for (int e = tid; e < 128 * d; e += 256) {
        int pt = e / d, k = e - pt * d;
        Xi_s[e] = X[(int64_t)(it * 128) * d + e];
        XjT[k * GRAD_XT_LD + pt] = X[(int64_t)(jt * 128) * d + e];
    }
    if (tid < d + 2) w_s[tid] = theta[tid];
    if (tid < 128) { ai_s[tid] = alpha[it * 128 + tid]; aj_s[tid] = alpha[jt * 128 + tid]; }
    __syncthreads();

    const int ty = tid >> 4, tx = tid & 15;
    const double sigma2 = w_s[d], nugget = w_s[d + 1];
    double acc[DMAX + 2];
#pragma unroll
    for (int k = 0; k < DMAX + 2; k++) acc[k] = 0.0;
    double w[DMAX];
#pragma unroll
    for (int k = 0; k < DMAX; k++) w[k] = (k < d) ? w_s[k] : 0.0;

#pragma unroll 1
    for (int a = 0; a < 8; a++) {
        const int li = ty * 8 + a, gi = it * 128 + li;
        double xi[DMAX];
#pragma unroll
        for (int k = 0; k < DMAX; k++) xi[k] = (k < d) ? Xi_s[li * d + k] : 0.0;
        const double ai = ai_s[li];
        const double* krow = Kinv + (int64_t)gi * ld + jt * 128;
#pragma unroll 1
        for (int b0 = 0; b0 < 8; b0 += 4) {
            double gk[4][DMAX], r2[4], wv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int lj = tx + 16 * (b0 + u), gj = jt * 128 + lj;
                const bool valid = (gj < gi) && (gi < N);            // strictly lower triangle, no padding (gj < gi < N)
                wv[u] = valid ? (__ldg(krow + lj) - ai * aj_s[lj]) * sigma2 : 0.0;
                r2[u] = 0.0;
            }
#pragma unroll
            for (int k = 0; k < DMAX; k++) {
                if (k < d) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const double df = xi[k] - XjT[k * GRAD_XT_LD + tx + 16 * (b0 + u)];
                        gk[u][k] = w[k] * df * df;
                        r2[u] += gk[u][k];
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; u++) gk[u][k] = 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                double cv, dv;
                if (KID == EXQ_KERNEL_MATERN52) {
                    const double s5 = sqrt_nonneg(5.0 * r2[u]), e = exp_nonpos(-s5);
                    cv = (1.0 + s5 + (5.0 / 3.0) * r2[u]) * e;
                    dv = (-5.0 / 6.0) * (1.0 + s5) * e;
                } else {
                    cv = exp_nonpos(-0.5 * r2[u]);
                    dv = -0.5 * cv;
                }
                const double f = wv[u] * dv;
#pragma unroll
                for (int k = 0; k < DMAX; k++) acc[k] = fma(f, gk[u][k], acc[k]);
                acc[DMAX] = fma(wv[u], cv, acc[DMAX]);
            }
        }
    }
    // diagonal entries: dK_ii/d ln sigma2 = sigma2, dK_ii/d ln nugget = nugget, no length-scale dependence
    if (it == jt && tid < 128) {
        const int gi = it * 128 + tid;
        if (gi < N) {
            const double W = Kinv[(int64_t)gi * ld + gi] - ai_s[tid] * ai_s[tid];
            acc[DMAX] += 0.5 * W * sigma2;
            acc[DMAX + 1] += 0.5 * W * nugget;
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int k = 0; k < DMAX + 2; k++) {
        double v = warp_sum(acc[k]);
        if (lane == 0) red[warp * (DMAX + 2) + k] = v;
    }
    __syncthreads();
    if (tid < DMAX + 2) {
        double sum = 0.0;
        for (int wv2 = 0; wv2 < 8; wv2++) sum += red[wv2 * (DMAX + 2) + tid];
        int tile = it * ntiles_side + jt;
        partial[((int64_t)blockIdx.z * ntiles_side * ntiles_side + tile) * (DMAX + 2) + tid] = sum;
    }
}

// sum the per-tile partials in fixed order -> grad[b][0..d-1], grad[b][d], grad[b][d+1]
__global__ void grad_reduce_kernel(const double* __restrict__ partial, int ntiles_side, int dmax, int d,
                                   double* __restrict__ grad /*[B][d+2]*/) {
    const int b = blockIdx.x;
    const int k = threadIdx.x;   // 0..d+1
    if (k >= d + 2) return;
    const int src = (k < d) ? k : (dmax + (k - d));
    double s = 0.0;
    for (int it = 0; it < ntiles_side; it++)
        for (int jt = 0; jt <= it; jt++)
            s += partial[((int64_t)b * ntiles_side * ntiles_side + it * ntiles_side + jt) * (dmax + 2) + src];
    grad[(int64_t)b * (d + 2) + k] = s;
}

// record the current best candidate as chosen[step] and stage its coordinates as x_new
// (an arg max over values that are all NaN leaves the sentinel index: nothing is read through it; the host reports it)
__global__ void select_best_kernel(const ArgPair* __restrict__ best, const double* __restrict__ Xc, long long M, int d,
                                   double* __restrict__ xnew, ArgPair* __restrict__ chosen, int step) {
    long long bi = best->i;
    if (bi >= 0 && bi < M)
        for (int k = threadIdx.x; k < d; k += blockDim.x) xnew[k] = Xc[bi * d + k];
    if (threadIdx.x == 0) chosen[step] = *best;
}

// explicit LOO refits: k_b[j] = sigma2 * corr(x_skip[b], Xb[b][j]) for j < N-1, 0 on the padding
__global__ void __launch_bounds__(256) loo_cross_vec_kernel(const double* __restrict__ X, const double* __restrict__ Xb,
                                                            int64_t strideX, const int* __restrict__ skip,
                                                            const double* __restrict__ theta, int64_t strideTheta,
                                                            int N, int d, int kernel_id, double* __restrict__ kvec,
                                                            int64_t strideK, int Np_out) {
    const int b = blockIdx.y;
    const double* th = theta + (int64_t)b * strideTheta;
    const double* xs = X + (int64_t)skip[b] * d;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Np_out; j += gridDim.x * blockDim.x) {
        double v = 0.0;
        if (j < N - 1) v = th[d] * point_corr(kernel_id, xs, Xb + (int64_t)b * strideX + (int64_t)j * d, th, d);
        kvec[(int64_t)b * strideK + j] = v;
    }
}

// mean = k^T alpha, var = max(sigma2 + nugget - |z|^2, 0) with z = Linv k, NES against y[skip]
__global__ void __launch_bounds__(256) loo_refit_finish_kernel(const double* __restrict__ kvec, const double* __restrict__ zvec,
                                                               const double* __restrict__ alpha, int64_t strideV,
                                                               const double* __restrict__ theta, int64_t strideTheta, int d,
                                                               const double* __restrict__ y, const int* __restrict__ skip,
                                                               int Np_out, double* __restrict__ mean, double* __restrict__ var,
                                                               double* __restrict__ nes) {
    __shared__ double r1[8], r2[8];
    const int b = blockIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int j = threadIdx.x; j < Np_out; j += 256) {
        s1 = fma(kvec[(int64_t)b * strideV + j], alpha[(int64_t)b * strideV + j], s1);
        double z = zvec[(int64_t)b * strideV + j];
        s2 = fma(z, z, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) { s1 += r1[k]; s2 += r2[k]; }
        const double* th = theta + (int64_t)b * strideTheta;
        double v = fmax(th[d] + th[d + 1] - s2, 0.0);
        mean[b] = s1;
        var[b] = v;
        nes[b] = nes_error(s1, v, y[skip[b]]);
    }
}

// gather rows of X / y leaving one index out (explicit LOO refits): system b skips skip[b]
__global__ void gather_skip_kernel(const double* __restrict__ X, const double* __restrict__ y, int N, int d,
                                   const int* __restrict__ skip, double* __restrict__ Xo, int64_t strideX,
                                   double* __restrict__ yo, int64_t strideY, int Np_out) {
    const int b = blockIdx.y;
    const int sk = skip[b];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Np_out; i += gridDim.x * blockDim.x) {
        int src = (i < sk) ? i : i + 1;
        bool valid = i < N - 1;
        for (int k = 0; k < d; k++) Xo[(int64_t)b * strideX + (int64_t)i * d + k] = valid ? X[(int64_t)src * d + k] : 0.0;
        yo[(int64_t)b * strideY + i] = valid ? y[src] : 0.0;
    }
}

}  // namespace exq
