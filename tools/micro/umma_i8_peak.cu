// Measured int8 tensor-core peak of this B200 for the roofline denominator of the tcgen05 kernels
// ("the int8 denominator is assumed, not measured").
//
//  (1) UTCIMMA issue-rate loop: one persistent CTA per SM, operands resident in shared memory (random bytes,
//      SWIZZLE_64B K-major tiles exactly as csrc/ozaki_i8.cuh lays them out), one thread issues
//      tcgen05.mma.cta_group::1.kind::i8 M=128, K=32, N in {64, 128, 256} back to back into TMEM, one
//      tcgen05.commit per 64 instructions.  No TMA, no epilogue: what the pipe can issue when nothing else is in the way.
//  (2) cuBLASLt IGEMM s8 x s8 -> s32, 8192^3 (TN), best of 10 and 2 s sustained: what a library kernel reaches
//      through the memory system.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/umma_i8_peak tools/micro/umma_i8_peak.cu -lcublasLt -lcuda
// run:   build/umma_i8_peak > gpurun_out/int8_peak.json
#include <cublasLt.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <vector>

#include "../../exauq-toolbox_b200/csrc/ozaki_i8.cuh"

using namespace exq;
using namespace exq::oz;

template <int N>
__global__ void __launch_bounds__(128, 1) issue_loop_kernel(int iters, unsigned seed) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    // A: 128 rows x 64 B, B: N rows x 64 B (two k steps of 32 each)
    const int bytes = (128 + N) * 64;
    unsigned x = seed + threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) {
        x = x * 1664525u + 1013904223u;
        smem[i] = (uint8_t)(x >> 24);
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        const uint64_t da = smem_desc_sw64(smem_u32(smem));
        const uint64_t db = smem_desc_sw64(smem_u32(smem + 128 * 64));
        uint32_t ph = 0;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 64; u++) {
                const int ks = u & 1;
                // alternate between two accumulator column blocks so that consecutive MMAs are independent
                const uint32_t d = tmem_base + (uint32_t)(((u >> 1) & 1) * (N < 256 ? N : 256));
                umma_i8(d, da + (uint64_t)((ks * 32) >> 4), db + (uint64_t)((ks * 32) >> 4), idesc_i8(128, N), (it | u) ? 1u : 0u);
            }
            umma_commit(&bar);
            mbar_wait_guard(&bar, ph);
            ph ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int N>
double run_issue_loop(int sms, double* sustained) {
    const int smem = (128 + N) * 64 + 2048;
    cudaFuncSetAttribute(issue_loop_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4000;
    const double ops = 2.0 * 128 * N * 32 * 64.0 * iters * sms;
    double best = 0.0;
    issue_loop_kernel<N><<<sms, 128, smem>>>(100, 1u);
    cudaDeviceSynchronize();
    for (int r = 0; r < 10; r++) {
        cudaEventRecord(e0);
        issue_loop_kernel<N><<<sms, 128, smem>>>(iters, 7u + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::max(best, ops / (ms * 1e-3) / 1e12);
    }
    // sustained: back to back for ~2 s
    cudaEventRecord(e0);
    int launches = 0;
    float ms = 0.f;
    do {
        for (int r = 0; r < 20; r++) issue_loop_kernel<N><<<sms, 128, smem>>>(iters, 99u + r);
        launches += 20;
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    } while (ms < 2000.f);
    *sustained = ops * launches / (ms * 1e-3) / 1e12;
    return best;
}

static bool cublaslt_igemm(int n, double* burst, double* sustained, char* err, size_t errlen) {
    cublasLtHandle_t lt;
    if (cublasLtCreate(&lt) != CUBLAS_STATUS_SUCCESS) { snprintf(err, errlen, "cublasLtCreate failed"); return false; }
    int8_t *A, *B;
    int32_t* C;
    cudaMalloc(&A, (size_t)n * n);
    cudaMalloc(&B, (size_t)n * n);
    cudaMalloc(&C, (size_t)n * n * 4);
    std::vector<int8_t> h((size_t)n * n);
    unsigned x = 12345u;
    for (auto& v : h) { x = x * 1664525u + 1013904223u; v = (int8_t)(x >> 24); }
    cudaMemcpy(A, h.data(), h.size(), cudaMemcpyHostToDevice);
    for (auto& v : h) { x = x * 1664525u + 1013904223u; v = (int8_t)(x >> 24); }
    cudaMemcpy(B, h.data(), h.size(), cudaMemcpyHostToDevice);
    cublasLtMatmulDesc_t op;
    cublasLtMatrixLayout_t la, lb, lc;
    cublasLtMatmulDescCreate(&op, CUBLAS_COMPUTE_32I, CUDA_R_32I);
    cublasOperation_t T = CUBLAS_OP_T, Nn = CUBLAS_OP_N;
    cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &T, sizeof(T));
    cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &Nn, sizeof(Nn));
    cublasLtMatrixLayoutCreate(&la, CUDA_R_8I, n, n, n);
    cublasLtMatrixLayoutCreate(&lb, CUDA_R_8I, n, n, n);
    cublasLtMatrixLayoutCreate(&lc, CUDA_R_32I, n, n, n);
    cublasLtMatmulPreference_t pref;
    cublasLtMatmulPreferenceCreate(&pref);
    size_t ws = (size_t)256 << 20;
    void* wsp;
    cudaMalloc(&wsp, ws);
    cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws, sizeof(ws));
    cublasLtMatmulHeuristicResult_t heur[8];
    int found = 0;
    cublasStatus_t st = cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, lc, lc, pref, 8, heur, &found);
    if (st != CUBLAS_STATUS_SUCCESS || found == 0) {
        snprintf(err, errlen, "cublasLtMatmulAlgoGetHeuristic: status %d, %d algorithms", (int)st, found);
        return false;
    }
    int32_t alpha = 1, beta = 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double ops = 2.0 * n * (double)n * n;
    *burst = 0.0;
    for (int a = 0; a < found; a++) {
        st = cublasLtMatmul(lt, op, &alpha, A, la, B, lb, &beta, C, lc, C, lc, &heur[a].algo, wsp, ws, 0);
        if (st != CUBLAS_STATUS_SUCCESS) continue;
        cudaDeviceSynchronize();
        for (int r = 0; r < 10; r++) {
            cudaEventRecord(e0);
            cublasLtMatmul(lt, op, &alpha, A, la, B, lb, &beta, C, lc, C, lc, &heur[a].algo, wsp, ws, 0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            *burst = std::max(*burst, ops / (ms * 1e-3) / 1e12);
        }
    }
    if (*burst == 0.0) { snprintf(err, errlen, "no cublasLt algorithm ran (last status %d)", (int)st); return false; }
    cudaEventRecord(e0);
    int runs = 0;
    float ms = 0.f;
    do {
        for (int r = 0; r < 20; r++) cublasLtMatmul(lt, op, &alpha, A, la, B, lb, &beta, C, lc, C, lc, &heur[0].algo, wsp, ws, 0);
        runs += 20;
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    } while (ms < 2000.f);
    *sustained = ops * runs / (ms * 1e-3) / 1e12;
    return true;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double s64, s128, s256;
    const double b64 = run_issue_loop<64>(sms, &s64);
    const double b128 = run_issue_loop<128>(sms, &s128);
    const double b256 = run_issue_loop<256>(sms, &s256);
    double lb = 0.0, ls = 0.0;
    char err[256] = "";
    const bool ok = cublaslt_igemm(8192, &lb, &ls, err, sizeof(err));
    cudaError_t ce = cudaGetLastError();
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"umma_i8_issue_loop_tops\": {\"n64\": {\"burst\": %.1f, \"sustained\": %.1f}, "
           "\"n128\": {\"burst\": %.1f, \"sustained\": %.1f}, \"n256\": {\"burst\": %.1f, \"sustained\": %.1f}}, "
           "\"cublaslt_igemm_8192_tops\": %s%s%s, \"cuda_error\": \"%s\", "
           "\"how\": \"tools/micro/umma_i8_peak.cu: tcgen05.mma.kind::i8 M=128 K=32 N=64/128/256 issued back to back by one thread per SM from "
           "resident SWIZZLE_64B shared-memory tiles of random bytes (best of 10 x 256000 MMAs per SM; sustained = 2 s back to back); "
           "cuBLASLt s8 x s8 -> s32 TN 8192^3 (best of 10 over the heuristic's algorithms; sustained = 2 s)\"}\n",
           prop.name, sms, b64, s64, b128, s128, b256, s256,
           ok ? "{\"burst\": " : "null", ok ? (std::to_string(lb) + ", \"sustained\": " + std::to_string(ls) + "}").c_str() : "",
           ok ? "" : (std::string(", \"cublaslt_error\": \"") + err + "\"").c_str(), cudaGetErrorString(ce));
    return 0;
}
