// DMMA.8x8x4 issue-rate microbenchmark: NW warps per CTA, 1 CTA per SM, ACC independent accumulators.
#include <cuda_runtime.h>
#include <cstdio>
template <int ACC>
__global__ void dmma_loop(double* out, int iters) {
    double c[ACC][2];
#pragma unroll
    for (int i = 0; i < ACC; i++) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ACC>
void run(int nw, int sms) {
    double* out; cudaMalloc(&out, sizeof(double) * sms * nw * 32 * 4);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dmma_loop<ACC><<<sms, nw * 32>>>(out, 100);
    cudaEventRecord(e0);
    dmma_loop<ACC><<<sms, nw * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 512.0 * ACC * iters * nw * sms;
    printf("warps/CTA=%2d acc=%2d: %.2f TFLOP/s\n", nw, ACC, flops / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    for (int nw : {4, 8, 16, 32}) { run<4>(nw, sms); run<16>(nw, sms); run<32>(nw, sms); }
    return 0;
}
