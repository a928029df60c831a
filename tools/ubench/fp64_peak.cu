// Micro-benchmark: FP64 peaks of the device -- DFMA (CUDA cores) and DMMA (mma.sync.m8n8k4.f64, the only FP64 tensor
// path; tcgen05 has no FP64) -- measured with independent accumulation chains per thread, all SMs busy, CUDA events.
// BASELINE.md section 2 / SURVEY.md 8(d) require these before any "% of FP64 pipe" claim.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double *out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void dmma_kernel(double *out, int iters, double a, double b) {
    double c0[CHAINS], c1[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) c0[c] = c1[c] = threadIdx.x + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[c]), "+d"(c1[c])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += c0[c] + c1[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int dev = 0, n_sm = 0, clk = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const int threads = 512, blocks = n_sm * 2, iters = 1 << 15;
    double *out;
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto report = [&](const char *name, double flops, float ms) {
        printf("{\"bench\": \"fp64_peak\", \"kind\": \"%s\", \"tflops\": %.3f, \"ms\": %.3f, \"n_sm\": %d, \"clock_khz_max\": %d}\n", name,
               flops / (ms * 1e-3) / 1e12, ms, n_sm, clk);
    };
    for (int rep = 0; rep < 3; ++rep) {
        float ms;
        cudaEventRecord(e0);
        dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) report("dfma (8 chains/thread)", 2.0 * 8 * (double)iters * threads * blocks, ms);
        cudaEventRecord(e0);
        dmma_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        // one m8n8k4 per warp = 8*8*4 multiply-adds
        if (rep) report("dmma m8n8k4 (4 chains/warp)", 2.0 * 256 * 4 * (double)iters * (threads / 32) * blocks, ms);
        cudaEventRecord(e0);
        dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) report("dmma m8n8k4 (8 chains/warp)", 2.0 * 256 * 8 * (double)iters * (threads / 32) * blocks, ms);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
