// Micro-benchmark: L1 cost of scattered 256-bit loads (ld.global.nc.v4.f64) as a function of how the 32 lanes of a warp
// spread over 128-byte lines and 32-byte sectors.  The table is L1-resident (16 KB), so the time per load instruction is
// the LSU / L1 data-stage cost of the access pattern.  Prints cycles per warp-level load instruction for each pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lsu_patterns lsu_patterns.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct G4 { double x, y, z, w; };
__device__ __forceinline__ G4 ldg256(const double *p) {
    G4 g;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(g.x), "=d"(g.y), "=d"(g.z), "=d"(g.w) : "l"(p));
    return g;
}
__device__ __forceinline__ double2 ldg128(const double *p) {
    double2 g;
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(g.x), "=d"(g.y) : "l"(p));
    return g;
}

// pattern: byte offset of lane l at iteration it inside a table of n_lines lines
__device__ __forceinline__ uint32_t offset_of(int pattern, int lane, int it, int n_lines) {
    uint32_t line, sector;
    switch (pattern) {
        case 0: line = (lane * 5 + it * 7) % n_lines; sector = 0; break;                     // 32 distinct lines
        case 1: line = ((lane >> 2) * 5 + it * 7) % n_lines; sector = lane & 3; break;       // 4 adjacent lanes: one line, 4 sectors
        case 2: line = ((lane >> 2) * 5 + it * 7) % n_lines; sector = 0; break;              // 4 adjacent lanes: same sector
        case 3: line = ((lane & 7) * 5 + it * 7) % n_lines; sector = lane >> 3; break;       // lanes l, l+8, l+16, l+24 share a line
        case 4: line = ((lane >> 1) * 5 + it * 7) % n_lines; sector = (lane & 1) * 2; break; // pairs share a line
        case 5: line = ((lane >> 3) * 5 + it * 7) % n_lines; sector = lane & 3; break;       // 8 adjacent lanes per line
        default: line = (it * 7) % n_lines; sector = 0; break;                               // broadcast
    }
    return line * 128u + sector * 32u;
}

template <int WIDTH>
__global__ void kern(const double *tab, int n_lines, int pattern, int iters, double *out, long long *cycles) {
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    const char *base = reinterpret_cast<const char *>(tab);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        const uint32_t off = offset_of(pattern, lane, it + (threadIdx.x >> 5), n_lines);
        if (WIDTH == 32) {
            const G4 g = ldg256(reinterpret_cast<const double *>(base + off));
            acc += g.x + g.w;
        } else {
            const double2 g = ldg128(reinterpret_cast<const double *>(base + off));
            acc += g.x + g.y;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    const int n_lines = 128, iters = 4096, threads = 512, blocks = 148;
    double *tab, *out;
    long long *cyc;
    cudaMalloc(&tab, n_lines * 128);
    cudaMemset(tab, 0, n_lines * 128);
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    long long h[148];
    const char *names[] = {"32 distinct lines", "quads: one line, 4 sectors", "quads: one sector", "stride-8 lanes share a line",
                           "pairs share a line", "octets: one line", "broadcast"};
    for (int width = 32; width >= 16; width -= 16)
        for (int p = 0; p < 7; ++p) {
            for (int rep = 0; rep < 2; ++rep) {
                if (width == 32) kern<32><<<blocks, threads>>>(tab, n_lines, p, iters, out, cyc);
                else kern<16><<<blocks, threads>>>(tab, n_lines, p, iters, out, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int b = 0; b < blocks; ++b) avg += (double)h[b];
            avg /= blocks;
            // 16 warps per SM issue `iters` load instructions each
            printf("{\"bench\": \"lsu_patterns\", \"width_bytes\": %d, \"pattern\": \"%s\", \"cycles_per_warp_load\": %.3f}\n", width,
                   names[p], avg / (iters * (threads / 32.0)));
        }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
