// Stand-alone check and timing of the int8 tcgen05 (Ozaki-sliced) variance product against the
// FP64 DMMA kernel and a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/ozaki_test tools/micro/ozaki_test.cu
//   ozaki_test <Np> <Mc> <mode> [reps]      mode 0: Linv-like random lower-triangular; 1: identity; 2: ones (all-k sums)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../exauq-toolbox_b200/csrc/gemm_dmma.cuh"
#include "../../exauq-toolbox_b200/csrc/ozaki_i8.cuh"
using namespace exq;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

template <int S>
int run(int Np, int Mc, int mode, int reps, const std::vector<double>& hL, const std::vector<double>& hK, double sigma2,
        const double* dL, const double* dK, const double* dPref, const std::vector<double>& Pref) {
    const int nrb = Np / 64;
    int8_t *lp, *cp;
    int* rexp;
    double* P;
    CK(cudaMalloc(&lp, (size_t)S * Np * Np));
    CK(cudaMalloc(&cp, (size_t)S * Mc * Np));
    CK(cudaMalloc(&rexp, Np * sizeof(int)));
    // the int8 kernel writes TWO partials per 64-row block (one per 32-row half: P[2 j + half][cand])
    CK(cudaMalloc(&P, (size_t)2 * nrb * Mc * 8));
    CK(cudaMemset(P, 0xff, (size_t)2 * nrb * Mc * 8));
    CK(oz::set_attrs<S>());
    int sm = 0;
    CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
    const int fexp = oz::scale_exponent(sigma2);
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);

    oz::row_exponent_lower_kernel<<<(Np + 7) / 8, 256>>>(dL, Np, Np, rexp);
    oz::launch_slice<S>(dL, Np, Np, Np, rexp, 0, 1, lp, 0);
    CK(cudaDeviceSynchronize());
    float ms_slice = 0, ms_main = 0;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(e0);
        oz::launch_slice<S>(dK, Np, Mc, Np, nullptr, fexp, 0, cp, 0);
        cudaEventRecord(e1);
        int rc = oz::launch_colsumsq<S>(cp, Mc, fexp, lp, rexp, Np, P, Mc, sm, 0);
        if (rc) { printf("launch_colsumsq rc=%d\n", rc); return 2; }
        cudaEventRecord(e2);
        CK(cudaEventSynchronize(e2));
        CK(cudaGetLastError());
        if (r > 0) {
            float a, b;
            cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
            ms_slice += a; ms_main += b;
        }
    }
    ms_slice /= reps; ms_main /= reps;
    std::vector<double> hP2((size_t)2 * nrb * Mc), hP((size_t)nrb * Mc);
    CK(cudaMemcpy(hP2.data(), P, hP2.size() * 8, cudaMemcpyDeviceToHost));
    for (int j = 0; j < nrb; j++)
        for (int c = 0; c < Mc; c++) hP[(size_t)j * Mc + c] = hP2[(size_t)(2 * j) * Mc + c] + hP2[(size_t)(2 * j + 1) * Mc + c];
    // compare per-candidate totals (what the variance uses) and per-block partials
    double max_rel_tot = 0, max_abs_tot = 0, max_rel_blk = 0;
    int bad = 0;
    for (int c = 0; c < Mc; c++) {
        double t = 0, tr = 0;
        for (int j = 0; j < nrb; j++) {
            double v = hP[(size_t)j * Mc + c], r = Pref[(size_t)j * Mc + c];
            t += v; tr += r;
            double rel = fabs(v - r) / fmax(fabs(r), 1e-300);
            if (fabs(r) > 1e-12 * sigma2 && rel > max_rel_blk) max_rel_blk = rel;
            if (!(rel < 1e-6) && fabs(r) > 1e-12 * sigma2 && bad < 6) {
                printf("  mismatch S=%d block %d cand %d: got %.17g want %.17g\n", S, j, c, v, r);
                bad++;
            }
        }
        max_abs_tot = fmax(max_abs_tot, fabs(t - tr));
        max_rel_tot = fmax(max_rel_tot, fabs(t - tr) / fmax(fabs(tr), 1e-300));
    }
    const double flops = (double)Np * Np * Mc;                         // FP64-equivalent (triangular k range)
    const double iops = flops * (S * (S + 1) / 2) * (1.0 + 1.0 / (Np / 64));   // int8 ops incl. the diagonal blocks' zero half
    printf("S=%d Np=%d Mc=%d mode=%d: total max abs %.3e max rel %.3e | block max rel %.3e | slice %.3f ms, mma %.3f ms "
           "-> %.1f TFLOP/s fp64-equivalent (%.2f Pop/s int8)%s\n",
           S, Np, Mc, mode, max_abs_tot, max_rel_tot, max_rel_blk, ms_slice, ms_main, flops / (ms_main * 1e-3) / 1e12,
           iops / (ms_main * 1e-3) / 1e15, bad ? "  ** MISMATCH **" : "");
    cudaFree(lp); cudaFree(cp); cudaFree(rexp); cudaFree(P);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    const int Np = argc > 1 ? atoi(argv[1]) : 256, Mc = argc > 2 ? atoi(argv[2]) : 256;
    const int mode = argc > 3 ? atoi(argv[3]) : 0, reps = argc > 4 ? atoi(argv[4]) : 3;
    const int smask = argc > 5 ? atoi(argv[5]) : 7;   // bit0: S=5, bit1: S=6, bit2: S=7
    const double sigma2 = 1.7;
    std::mt19937_64 rng(1234);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::vector<double> hL((size_t)Np * Np, 0.0), hK((size_t)Mc * Np);
    for (int i = 0; i < Np; i++)
        for (int k = 0; k <= i; k++) {
            double v;
            if (mode == 1) v = (i == k) ? 1.0 : 0.0;
            else if (mode == 2) v = 1.0;
            else v = (i == k) ? 5.0 + U(rng) : U(rng) * exp(-0.002 * (i - k)) * pow(10.0, 2.0 * U(rng));
            hL[(size_t)i * Np + k] = v;
        }
    // scratch above the diagonal outside the 128-blocks must be ignored by the slicer
    for (int i = 0; i < Np; i++)
        for (int k = (i / 128 + 1) * 128; k < Np; k++) hL[(size_t)i * Np + k] = 1e300;
    for (size_t e = 0; e < hK.size(); e++) {
        double u = 0.5 * (U(rng) + 1.0);
        hK[e] = mode == 2 ? 1.0 : sigma2 * exp(-6.0 * u * u);
    }
    double *dL, *dK, *dPref;
    const int nrb = Np / 64;
    CK(cudaMalloc(&dL, hL.size() * 8)); CK(cudaMalloc(&dK, hK.size() * 8)); CK(cudaMalloc(&dPref, (size_t)nrb * Mc * 8));
    CK(cudaMemcpy(dL, hL.data(), hL.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dK, hK.data(), hK.size() * 8, cudaMemcpyHostToDevice));

    // FP64 reference on the device: the production DMMA kernel (64-row tiles -> same P layout)
    using G = GemmCfg<64>;
    CK(cudaFuncSetAttribute(gemm_dmma_kernel<LAYOUT_K, LAYOUT_K, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    GemmArgs a{};
    a.A = dL; a.B = dK; a.C = dPref; a.lda = Np; a.ldb = Np; a.ldc = Mc; a.M = Np; a.N = Mc; a.K = Np;
    a.krange = KR_LE_I; a.cmode = C_COLSUMSQ; a.raster = 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_ref = 0;
    for (int r = 0; r < 2; r++) {
        cudaEventRecord(e0);
        gemm_dmma_kernel<LAYOUT_K, LAYOUT_K, 64><<<dim3((Np / 64) * (Mc / 128), 1, 1), G::THREADS, G::SMEM_BYTES>>>(a);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms_ref, e0, e1);
    }
    CK(cudaGetLastError());
    std::vector<double> Pref((size_t)nrb * Mc);
    CK(cudaMemcpy(Pref.data(), dPref, Pref.size() * 8, cudaMemcpyDeviceToHost));
    printf("DMMA reference: %.3f ms -> %.1f TFLOP/s\n", ms_ref, (double)Np * Np * Mc / (ms_ref * 1e-3) / 1e12);
    if ((size_t)Np * Np * Mc <= (size_t)1 << 31) {   // host cross-check of the reference itself
        double worst = 0;
        for (int c = 0; c < Mc; c += 37)
            for (int j = 0; j < nrb; j++) {
                double s = 0;
                for (int i = j * 64; i < j * 64 + 64; i++) {
                    double z = 0;
                    for (int k = 0; k <= i; k++) z += hL[(size_t)i * Np + k] * hK[(size_t)c * Np + k];
                    s += z * z;
                }
                worst = fmax(worst, fabs(s - Pref[(size_t)j * Mc + c]) / fmax(fabs(s), 1e-300));
            }
        printf("host vs DMMA reference: max rel %.3e\n", worst);
    }
    int rc = 0;
    if (smask & 1) rc |= run<5>(Np, Mc, mode, reps, hL, hK, sigma2, dL, dK, dPref, Pref);
    if (smask & 2) rc |= run<6>(Np, Mc, mode, reps, hL, hK, sigma2, dL, dK, dPref, Pref);
    if (smask & 4) rc |= run<7>(Np, Mc, mode, reps, hL, hK, sigma2, dL, dK, dPref, Pref);
    return rc;
}
