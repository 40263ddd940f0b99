// Stand-alone check + timing of the 128 x 128 potrf+trtri leaf (per-phase clocks with -DEXQ_LEAF_PROFILE).
#include <cstdio>
#include <cmath>
#include <vector>
#include "../../exauq-toolbox_b200/csrc/leaf.cuh"
using namespace exq;
int main() {
    const int n = 128, B = 1;
    std::vector<double> A(n * n), L(n * n, 0.0);
    for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
        double r2 = 0; for (int k = 0; k < 3; k++) { double d = sin(1.3 * i + k) - sin(1.3 * j + k); r2 += d * d; }
        double a = sqrt(5 * r2 / 0.25); A[i * n + j] = A[j * n + i] = (1 + a + a * a / 3) * exp(-a) + (i == j ? 1e-6 : 0);
    }
    for (int j = 0; j < n; j++) {       // host Cholesky
        double s = A[j * n + j]; for (int k = 0; k < j; k++) s -= L[j * n + k] * L[j * n + k];
        L[j * n + j] = sqrt(s);
        for (int i = j + 1; i < n; i++) { double t = A[i * n + j]; for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k]; L[i * n + j] = t / L[j * n + j]; }
    }
    double *dA, *dL; int* info;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&info, 4); cudaMemset(info, 0, 4);
    cudaFuncSetAttribute(leaf_potrf_trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM_BYTES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 6; rep++) {
        cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        leaf_potrf_trtri_kernel<<<B, LEAF_THREADS, LEAF_SMEM_BYTES>>>(dA, dL, n, n * n, n * n, 0, info);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
    }
    std::vector<double> gL(n * n), gI(n * n);
    cudaMemcpy(gL.data(), dA, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(gI.data(), dL, n * n * 8, cudaMemcpyDeviceToHost);
    double eL = 0, eI = 0;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        eL = fmax(eL, fabs(gL[i * n + j] - (j <= i ? L[i * n + j] : 0.0)));
        double s = 0; for (int k = 0; k < n; k++) s += gI[i * n + k] * (k >= j ? L[k * n + j] : 0.0);   // Linv * L = I
        eI = fmax(eI, fabs(s - (i == j)));
    }
    int hinfo; cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
    printf("leaf: best %.1f us, max|L-Lref| %.2e, max|Linv L - I| %.2e, info %d (%s)\n", best * 1e3, eL, eI, hinfo, cudaGetErrorString(cudaGetLastError()));
#ifdef EXQ_LEAF_PROFILE
    long long prof[64]; cudaMemcpyFromSymbol(prof, leaf_prof, sizeof(prof));
    const char* names[] = {"load+setup", "1a diag chol", "1b row subst + diag inv", "1c rank-8 update", "2.1 finalise row", "2.2 update", "writeback"};
    for (int i = 0; i < 7; i++) printf("  %-26s %8.0f cycles/call\n", names[i], prof[i] / 6.0);
#endif
    return 0;
}
