// Stand-alone check + per-step timeline of the fused sub-tree kernel (fused_node.cuh) on one 512- or 1024-row block.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DEXQ_FUSED_PROFILE -o build/fused_bench tools/micro/fused_bench.cu
//   build/fused_bench [n=512] [B=1] [G=8]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../exauq-toolbox_b200/csrc/fused_node.cuh"
using namespace exq;
int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 512, B = argc > 2 ? atoi(argv[2]) : 1, G = argc > 3 ? atoi(argv[3]) : 8;
    std::vector<double> A((size_t)n * n), L((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
        double r2 = 0; for (int k = 0; k < 4; k++) { double d = sin(1.3 * i + k) - sin(1.3 * j + k); r2 += d * d; }
        double a = sqrt(5 * r2 / 0.09); A[(size_t)i * n + j] = A[(size_t)j * n + i] = (1 + a + a * a / 3) * exp(-a) + (i == j ? 1e-6 : 0);
    }
    for (int j = 0; j < n; j++) {       // host Cholesky
        double s = A[(size_t)j * n + j]; for (int k = 0; k < j; k++) s -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
        L[(size_t)j * n + j] = sqrt(s);
        for (int i = j + 1; i < n; i++) { double t = A[(size_t)i * n + j]; for (int k = 0; k < j; k++) t -= L[(size_t)i * n + k] * L[(size_t)j * n + k]; L[(size_t)i * n + j] = t / L[(size_t)j * n + j]; }
    }
    double *dA, *dL; int* info;
    const size_t sq = (size_t)n * n;
    cudaMalloc(&dA, sq * 8 * B); cudaMalloc(&dL, sq * 8 * B); cudaMalloc(&info, 4 * B); cudaMemset(info, 0, 4 * B);
    cudaFuncSetAttribute(fused_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM_BYTES);
    if (G > 8) cudaFuncSetAttribute(fused_factor_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    FusedArgs a{};
    a.A = dA; a.Linv = dL; a.ld = n; a.stride = sq; a.info = info; a.n_ops = 0;
    if (!fused_build(a.ops, a.n_ops, 0, n, G)) { printf("step list overflow\n"); return 1; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G, B, 1); cfg.blockDim = dim3(FUSED_THREADS, 1, 1); cfg.dynamicSmemBytes = FUSED_SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 6; rep++) {
        for (int b = 0; b < B; b++) cudaMemcpy(dA + b * sq, A.data(), sq * 8, cudaMemcpyHostToDevice);
        cudaMemset(dL, 0, sq * 8 * B);
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchKernelEx(&cfg, fused_factor_kernel, a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
    }
    std::vector<double> gL(sq), gI(sq);
    cudaMemcpy(gL.data(), dA + (B - 1) * sq, sq * 8, cudaMemcpyDeviceToHost); cudaMemcpy(gI.data(), dL + (B - 1) * sq, sq * 8, cudaMemcpyDeviceToHost);
    double eL = 0, eI = 0;
    for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
        if ((i / 128) == (j / 128)) eL = fmax(eL, fabs(gL[(size_t)i * n + j] - L[(size_t)i * n + j]));   // L lives on the diagonal 128-blocks
        double s = 0; for (int k = j; k <= i; k++) s += gI[(size_t)i * n + k] * L[(size_t)k * n + j];   // Linv * L = I
        eI = fmax(eI, fabs(s - (i == j)));
    }
    int hinfo; cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
    printf("fused n=%d B=%d G=%d: best %.1f us, %d steps, max|L-Lref| %.2e, max|Linv L - I| %.2e, info %d (%s)\n", n, B, G, best * 1e3,
           a.n_ops, eL, eI, hinfo, cudaGetErrorString(cudaGetLastError()));
#ifdef EXQ_FUSED_PROFILE
    static unsigned long long st[16][48][2];
    cudaMemcpyFromSymbol(st, fused_stamps, sizeof(st));
    const unsigned long long t0 = st[0][a.n_ops][0];
    printf("  step type          M    N    K tile sync | rank0 work-done  barrier-passed | slowest rank work-done  (us since kernel start)\n");
    for (int i = 0; i < a.n_ops; i++) {
        const FusedOp& o = a.ops[i];
        unsigned long long w = 0; for (int r = 0; r < G; r++) w = st[r][i][0] > w ? st[r][i][0] : w;
        printf("  %3d  %-9s %4d %4d %4d   %d   %d    | %8.2f %8.2f | %8.2f\n", i, o.type == FOP_LEAF ? "leaf" : (o.skip_rank0 ? "gemm||leaf" : "gemm"),
               o.M, o.N, o.K, o.tile, o.sync_after, (st[0][i][0] - t0) * 1e-3, (st[0][i][1] - t0) * 1e-3, (w - t0) * 1e-3);
    }
#endif
    return 0;
}
