// L2 -> SM bandwidth of this B200, the feed rate the int8 tcgen05 kernels depend on: per 64-byte k block a CTA pulls
// S x (128 + 64) rows x 64 B of digit planes (86 KB at S = 7) for S(S+1)/2 x 1 Mop of MMAs, i.e. 2.9e-3 B per int8 op
// at S = 7 and 3.4e-3 at S = 6 -- at the chip's int8 issue rate that is 10.8 / 12.3 TB/s out of L2.
// One persistent CTA per SM streams an L2-resident buffer (default 48 MB; every CTA reads all of it, like the tiles of a
// GEMM wave share their operand planes) into shared memory with cp.async.bulk (TMA, the path the kernels use), 4 stages
// of 32 KB in flight, and with plain 16-byte ld.global.cg loads for comparison.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/l2_bw tools/micro/l2_bw.cu
// run:   build/l2_bw > gpurun_out/l2_bw.json
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#include "../../exauq-toolbox_b200/csrc/exq_common.cuh"

using namespace exq;

constexpr int CHUNK = 32 * 1024, STAGES = 4;

__global__ void __launch_bounds__(128, 1) tma_stream_kernel(const uint8_t* __restrict__ buf, size_t bytes, int passes,
                                                            unsigned long long* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const size_t n_chunks = bytes / CHUNK;
    const size_t total = n_chunks * (size_t)passes;
    unsigned long long acc = 0;
    if (threadIdx.x == 0) {
        // every CTA starts at its own offset so that the SMs do not all hit the same L2 slice at once
        size_t issued = 0, done = 0;
        const size_t off = (size_t)blockIdx.x * 977 % n_chunks;
        uint32_t ph[STAGES] = {0, 0, 0, 0};
        for (; issued < total && issued < STAGES; issued++) {
            mbar_expect_tx(&full[issued % STAGES], CHUNK);
            tma_bulk_g2s(smem + (issued % STAGES) * CHUNK, buf + ((issued + off) % n_chunks) * CHUNK, CHUNK, &full[issued % STAGES]);
        }
        for (; done < total; done++) {
            const int s = (int)(done % STAGES);
            mbar_wait(&full[s], ph[s]);
            ph[s] ^= 1u;
            acc += smem[s * CHUNK + (done & 1023)];
            if (issued < total) {
                mbar_expect_tx(&full[s], CHUNK);
                tma_bulk_g2s(smem + s * CHUNK, buf + ((issued + off) % n_chunks) * CHUNK, CHUNK, &full[s]);
                issued++;
            }
        }
        sink[blockIdx.x] = acc;
    }
}

__global__ void __launch_bounds__(512, 1) ldg_stream_kernel(const uint4* __restrict__ buf, size_t n16, int passes,
                                                            unsigned long long* sink) {
    unsigned long long acc = 0;
    const size_t off = ((size_t)blockIdx.x * 977 * 2048) % n16;
    for (int p = 0; p < passes; p++) {
        for (size_t i = threadIdx.x; i < n16; i += 512 * 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                size_t j = i + (size_t)u * 512 + off;
                if (j >= n16) j -= n16;
                v[u] = __ldcg(buf + j);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
    }
    if (threadIdx.x == 0) sink[blockIdx.x] = acc;
    if (acc == 0x1234567ull) sink[blockIdx.x + 1] = acc;
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? (size_t)atoi(argv[1]) : 48;
    const size_t bytes = mb << 20;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint8_t* buf;
    unsigned long long* sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, (sms + 1) * sizeof(unsigned long long));
    cudaMemset(buf, 1, bytes);
    cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int passes = 20;
    double best_tma = 0, best_ldg = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        tma_stream_kernel<<<sms, 128, STAGES * CHUNK>>>(buf, bytes, passes, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best_tma = std::max(best_tma, (double)sms * bytes * passes / (ms * 1e-3) / 1e12);
        cudaEventRecord(e0);
        ldg_stream_kernel<<<sms, 512>>>((const uint4*)buf, bytes / 16, passes, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best_ldg = std::max(best_ldg, (double)sms * bytes * passes / (ms * 1e-3) / 1e12);
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"buffer_mb\": %zu, \"l2_to_sm_tbs\": {\"tma_bulk_4x32KB_in_flight\": %.3f, "
           "\"ld_global_cg_16B\": %.3f}, \"cuda_error\": \"%s\", \"how\": \"tools/micro/l2_bw.cu: one CTA per SM streams the "
           "same L2-resident buffer %d times (best of 5)\"}\n",
           prop.name, sms, mb, best_tma, best_ldg, cudaGetErrorString(cudaGetLastError()), passes);
    return 0;
}
