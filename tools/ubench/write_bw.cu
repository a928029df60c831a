// Micro-benchmark: WRITE bandwidth of the device.  MEASURED_PEAKS.json's hbm_gbs is a copy (half reads, half writes); the
// row-item assembly kernel is write-dominated (2.73 GB written, 0.6 GB read per launch), so its practical ceiling is the
// write-only rate, measured here three ways: cudaMemsetAsync, a grid-stride 16-byte store kernel, and bulk copies
// shared -> global (cp.async.bulk, the path the assembly kernel's tiles take: 148 persistent CTAs, 4 tiles of 48 KB).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void store_kernel(double2 *out, size_t n) {
    const double2 v = make_double2(1.0, 2.0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}

__global__ void read_kernel(const double2 *in, size_t n, double *sink) {
    double acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = in[i];
        acc += v.x + v.y;
    }
    if (acc == 12345.678) *sink = acc;
}

// every CTA streams its contiguous share of `out` as bulk copies of `tile_bytes` from NBUF shared-memory tiles
template <int NBUF>
__global__ void __launch_bounds__(128, 1) bulk_store_kernel(char *out, size_t bytes, int tile_bytes) {
    extern __shared__ __align__(128) char smem[];
    for (int i = threadIdx.x; i < NBUF * tile_bytes / 8; i += blockDim.x) reinterpret_cast<double *>(smem)[i] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t n_tiles = bytes / tile_bytes;
    const size_t t0 = n_tiles * blockIdx.x / gridDim.x, t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
    int b = 0;
    for (size_t t = t0; t < t1; ++t) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem + (size_t)b * tile_bytes);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * tile_bytes), "r"(src),
                     "r"(tile_bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
        b = (b + 1) % NBUF;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const size_t bytes = (size_t)2800 << 20;
    char *buf, *buf2;
    double *sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&buf2, bytes);
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto timed = [&](const char *name, auto fn, double moved) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            fn();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("{\"bench\": \"write_bw\", \"kind\": \"%s\", \"gbytes_per_s\": %.1f, \"ms\": %.4f, \"gbytes\": %.3f}\n", name,
               moved / 1e9 / (best * 1e-3), best, moved / 1e9);
    };
    timed("cudaMemsetAsync", [&] { cudaMemsetAsync(buf, 0, bytes); }, (double)bytes);
    timed("st.global.v2.f64 grid-stride, 148x8 CTAs x 256", [&] { store_kernel<<<148 * 8, 256>>>((double2 *)buf, bytes / 16); },
          (double)bytes);
    timed("ld.global.v2.f64 grid-stride (read only)", [&] { read_kernel<<<148 * 8, 256>>>((const double2 *)buf, bytes / 16, sink); },
          (double)bytes);
    timed("cudaMemcpyAsync d2d (read + write)", [&] { cudaMemcpyAsync(buf2, buf, bytes, cudaMemcpyDeviceToDevice); }, 2.0 * bytes);
    const int tile = 48 * 1024;
    cudaFuncSetAttribute(bulk_store_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * tile);
    timed("cp.async.bulk shared->global, 148 CTAs, 4 x 48 KB tiles", [&] { bulk_store_kernel<4><<<148, 128, 4 * tile>>>(buf, bytes, tile); },
          (double)bytes);
    cudaFuncSetAttribute(bulk_store_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * tile);
    timed("cp.async.bulk shared->global, 148 CTAs, 2 x 48 KB tiles", [&] { bulk_store_kernel<2><<<148, 128, 2 * tile>>>(buf, bytes, tile); },
          (double)bytes);
    // write stream + concurrent read stream (the assembly kernel's mix: 2.73 GB written, 0.63 GB read)
    cudaStream_t s2;
    cudaStreamCreate(&s2);
    timed("bulk store 2.8 GB with a concurrent 0.65 GB read stream", [&] {
        read_kernel<<<148, 256, 0, s2>>>((const double2 *)buf2, ((size_t)650 << 20) / 16, sink);
        bulk_store_kernel<4><<<148, 128, 4 * tile>>>(buf, bytes, tile);
        cudaStreamSynchronize(s2);
    }, (double)bytes + (double)((size_t)650 << 20));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
