// Stand-alone check of the int8 tcgen05 tile GEMM (ozaki_gemm.cuh) in the five operand
// configurations the potrf+trtri recursion and LAUUM use, against a host FP64 reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/ozaki_gemm_test tools/micro/ozaki_gemm_test.cu
//   ozaki_gemm_test <n> <B> [time_n] [time_B]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../exauq-toolbox_b200/csrc/ozaki_gemm.cuh"
using namespace exq;

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

struct Case {
    const char* name;
    int transA, maskA, transB, maskB, krange, lower_only, cmode, same;
};

template <int S>
int run_case(const Case& c, int n, int B, int ld, const std::vector<double>& hA, const std::vector<double>& hB,
             const std::vector<double>& hC0, bool check, int sm) {
    const int64_t sq = (int64_t)ld * ld;
    double *dA, *dB, *dC;
    int8_t* planes;
    int* exps;
    CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dB, hB.size() * 8)); CK(cudaMalloc(&dC, hC0.size() * 8));
    CK(cudaMalloc(&planes, (size_t)2 * S * B * n * n)); CK(cudaMalloc(&exps, (size_t)2 * B * n * sizeof(int)));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, hC0.data(), hC0.size() * 8, cudaMemcpyHostToDevice));
    int8_t* pA = planes;
    int8_t* pB = c.same ? pA : planes + (size_t)S * B * n * n;
    int* eA = exps;
    int* eB = c.same ? eA : exps + (size_t)B * n;
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    oz::GArgs a{};
    a.expA = eA; a.expB = eB; a.C = dC; a.ldc = ld; a.strideC = sq; a.M = n; a.N = n; a.K = n; a.B = B;
    a.krange = c.krange; a.lower_only = c.lower_only; a.cmode = c.cmode;
    a.heavy_last = (c.krange == KR_LE_J || c.krange == KR_LE_I) ? 1 : 0;
    a.sys_slow = 1;
    uint32_t* dsched = nullptr;
    if (getenv("OZ_SCHED") && atoi(getenv("OZ_SCHED"))) {
        std::vector<uint32_t> list;
        const int items = (n / oz::BLK_C) * (n / oz::BLK_R) * B;
        const int grid = items < sm ? items : sm;
        oz::make_schedule(n, n, n, B + (getenv("OZ_SCHED_EXTRA") ? atoi(getenv("OZ_SCHED_EXTRA")) : 0), c.krange, c.lower_only, grid, list);
        CK(cudaMalloc(&dsched, list.size() * 4));
        CK(cudaMemcpy(dsched, list.data(), list.size() * 4, cudaMemcpyHostToDevice));
        a.sched = dsched; a.n_sched = (int)list.size(); a.sched_grid = grid;
    }
    float ms_s = 0, ms_g = 0;
    const int reps = check ? 1 : 3;
    for (int r = 0; r < reps; r++) {
        if (c.cmode == C_SUB) CK(cudaMemcpy(dC, hC0.data(), hC0.size() * 8, cudaMemcpyHostToDevice));
        cudaEventRecord(e0);
        oz::slice_operand<S>(dA, ld, sq, n, n, B, c.transA, c.maskA, eA, pA, 0);
        if (!c.same) oz::slice_operand<S>(dB, ld, sq, n, n, B, c.transB, c.maskB, eB, pB, 0);
        cudaEventRecord(e1);
        int rc = oz::launch_gemm<S>(pA, pB, a, sm, 0);
        if (rc) { printf("launch rc=%d\n", rc); return 2; }
        cudaEventRecord(e2);
        CK(cudaEventSynchronize(e2));
        CK(cudaGetLastError());
        cudaEventElapsedTime(&ms_s, e0, e1); cudaEventElapsedTime(&ms_g, e1, e2);
    }
    double fl = 2.0 * n * (double)n * n * B;
    if (c.krange != KR_FULL) fl *= 0.5;
    if (c.lower_only && c.krange == KR_FULL) fl *= 0.5;
    if (c.lower_only && c.krange != KR_FULL) fl *= 1.0 / 3.0;
    int bad = 0;
    double worst = 0;
    if (check) {
        std::vector<double> hC(hC0.size());
        CK(cudaMemcpy(hC.data(), dC, hC.size() * 8, cudaMemcpyDeviceToHost));
        auto opA = [&](int b, int i, int k) {
            bool keep = c.maskA == oz::MASK_NONE || (c.maskA == oz::MASK_LE ? k <= i : k >= i);
            if (!keep) return 0.0;
            return c.transA ? hA[b * sq + (int64_t)k * ld + i] : hA[b * sq + (int64_t)i * ld + k];
        };
        const std::vector<double>& srcB = c.same ? hA : hB;
        auto opB = [&](int b, int j, int k) {
            bool keep = c.maskB == oz::MASK_NONE || (c.maskB == oz::MASK_LE ? k <= j : k >= j);
            if (!keep) return 0.0;
            return c.transB ? srcB[b * sq + (int64_t)k * ld + j] : srcB[b * sq + (int64_t)j * ld + k];
        };
        for (int b = 0; b < B; b++)
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                    if (c.lower_only && (j / 64) * 64 > (i / 128) * 128 + 127) continue;   // tile not computed
                    // the digit scheme bounds the error by ~K 2^-8S max|a_i| max|b_j| (row-scaled fixed point)
                    double z = 0, ma = 0, mb = 0;
                    for (int k = 0; k < n; k++) {
                        const double x = opA(b, i, k), y = opB(b, j, k);
                        z += x * y; ma = fmax(ma, fabs(x)); mb = fmax(mb, fabs(y));
                    }
                    const double zabs = ma * mb;
                    double want = c.cmode == C_STORE ? z : (c.cmode == C_NEG ? -z : hC0[b * sq + (int64_t)i * ld + j] - z);
                    double got = hC[b * sq + (int64_t)i * ld + j];
                    double err = fabs(got - want) / fmax(zabs, 1e-300);
                    if (err > worst) worst = err;
                    if (!(err < 1e-11) && bad < 5) {
                        printf("   %s b=%d i=%d j=%d got %.17g want %.17g\n", c.name, b, i, j, got, want);
                        bad++;
                    }
                }
    }
    printf("S=%d %-28s n=%d B=%d: %s slice %.3f ms, gemm %.3f ms -> %.1f TFLOP/s fp64-equivalent", S, c.name, n, B,
           bad ? "** MISMATCH **" : "ok", ms_s, ms_g, fl / (ms_g * 1e-3) / 1e12);
    if (check) printf(", max err / (max|a_i| max|b_j|) = %.2e", worst);
    printf("\n");
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(planes); cudaFree(exps); cudaFree(dsched);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 256, B = argc > 2 ? atoi(argv[2]) : 2;
    const int tn = argc > 3 ? atoi(argv[3]) : 0;
    const int tB = argc > 4 ? atoi(argv[4]) : B;
    int sm = 0;
    CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
    CK(oz::set_gemm_attrs<6>());
    CK(oz::set_gemm_attrs<7>());
    const Case cases[] = {
        {"(1) A21 * Linv11^T   k<=j", 0, oz::MASK_NONE, 0, oz::MASK_LE, KR_LE_J, 0, C_STORE, 0},
        {"(2) A22 -= L21 L21^T lower", 0, oz::MASK_NONE, 0, oz::MASK_NONE, KR_FULL, 1, C_SUB, 1},
        {"(3) L21 * Linv11     k>=j", 0, oz::MASK_NONE, 1, oz::MASK_GE, KR_GE_J, 0, C_STORE, 0},
        {"(4) -Linv22 * T      k<=i", 0, oz::MASK_LE, 1, oz::MASK_NONE, KR_LE_I, 0, C_NEG, 0},
        {"(5) LAUUM Linv^T Linv", 1, oz::MASK_GE, 1, oz::MASK_GE, KR_GE_I, 1, C_STORE, 1},
    };
    int rc = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int nn = pass == 0 ? n : tn;
        const int B = pass == 0 ? atoi(argc > 2 ? argv[2] : "2") : tB;
        if (nn <= 0) continue;
        const int ld = nn + 128;   // sub-matrix inside a larger system
        std::mt19937_64 rng(7);
        std::uniform_real_distribution<double> U(-1.0, 1.0);
        std::vector<double> hA((size_t)B * ld * ld), hB(hA.size()), hC(hA.size());
        for (auto& v : hA) v = U(rng) * pow(10.0, 1.5 * U(rng));
        for (auto& v : hB) v = U(rng) * pow(10.0, 1.5 * U(rng));
        for (auto& v : hC) v = 100.0 * U(rng);
        for (const Case& c : cases) {
            rc |= run_case<7>(c, nn, B, ld, hA, hB, hC, pass == 0, sm);
            if (pass == 0) rc |= run_case<6>(c, nn, B, ld, hA, hB, hC, true, sm);
        }
    }
    return rc;
}
