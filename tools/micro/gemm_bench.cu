// Stand-alone timing of gemm_dmma_kernel in the variance-GEMM shape (Z = Linv * Kc^T, k <= i,
// column-sum-of-squares epilogue) and the SYRK shape, for both row-tile sizes.
#include <cstdio>
#include <vector>
#include "../../exauq-toolbox_b200/csrc/gemm_dmma.cuh"
using namespace exq;
template <int BM>
void run(int Np, int Mc, double* A, double* B, double* P, double* C) {
    using Cfg = GemmCfg<BM>;
    cudaFuncSetAttribute(gemm_dmma_kernel<LAYOUT_K, LAYOUT_K, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int shape = 0; shape < 2; shape++) {
        GemmArgs a{};
        double flops;
        dim3 grid;
        if (shape == 0) {
            a.A = A; a.B = B; a.C = P; a.lda = Np; a.ldb = Np; a.ldc = Mc; a.M = Np; a.N = Mc; a.K = Np;
            a.krange = KR_LE_I; a.cmode = C_COLSUMSQ; a.raster = 8;
            grid = dim3((Np / BM) * (Mc / 128), 1, 1);
            flops = (double)Np * Np * Mc;
        } else {
            a.A = A; a.B = A; a.C = C; a.lda = a.ldb = a.ldc = Np; a.M = a.N = Np; a.K = Np;
            a.krange = KR_FULL; a.lower_only = 1; a.cmode = C_SUB; a.raster = 0;
            grid = dim3(Np / 128, Np / BM, 1);
            flops = (double)Np * Np * Np;
        }
        auto launch = [&]() { gemm_dmma_kernel<LAYOUT_K, LAYOUT_K, BM><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES>>>(a); };
        for (int w = 0; w < 2; w++) launch();
        cudaEventRecord(e0);
        const int reps = 5;
        for (int r = 0; r < reps; r++) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("BM=%3d %s Np=%d Mc=%d: %.3f ms  %.2f TFLOP/s (%s)\n", BM, shape == 0 ? "variance" : "syrk    ", Np, Mc,
               ms / reps, flops * reps / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
}
int main(int argc, char** argv) {
    int Np = argc > 1 ? atoi(argv[1]) : 4096, Mc = argc > 2 ? atoi(argv[2]) : 32768;
    double *A, *B, *P, *C;
    cudaMalloc(&A, (size_t)Np * Np * 8); cudaMalloc(&B, (size_t)Mc * Np * 8);
    cudaMalloc(&P, (size_t)(Np / 64) * Mc * 8); cudaMalloc(&C, (size_t)Np * Np * 8);
    cudaMemset(A, 0, (size_t)Np * Np * 8); cudaMemset(B, 0, (size_t)Mc * Np * 8); cudaMemset(C, 0, (size_t)Np * Np * 8);
    run<128>(Np, Mc, A, B, P, C);
    run<64>(Np, Mc, A, B, P, C);
    return 0;
}
