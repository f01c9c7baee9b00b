// L2 -> SM gather bandwidth ceiling: every warp reads random rows of ROWB bytes (coalesced, 16 B per lane) from a table of
// `rows` rows; a table smaller than L2 measures the L2->SM path, a larger one the DRAM gather path.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2gather l2gather.cu && ./l2gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ROWB, bool NOALLOC>
__global__ void __launch_bounds__(512) gather(const uint4 *__restrict__ tab, uint32_t rows, int iters, float *out) {
    const int lane = threadIdx.x & 31;
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) / 32 * 2654435761u + 12345u;
    float acc = 0.f;
    constexpr int V = ROWB / 512;        // uint4 loads per lane and row
    for (int it = 0; it < iters; ++it) {
        uint4 v[8][V];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s = s * 1664525u + 1013904223u;
            const uint32_t r = (uint32_t)(((uint64_t)s * rows) >> 32);
            const uint4 *p = tab + (size_t)r * (ROWB / 16) + lane;
#pragma unroll
            for (int q = 0; q < V; ++q) {
                if (NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j][q].x), "=r"(v[j][q].y), "=r"(v[j][q].z), "=r"(v[j][q].w) : "l"(p + 32 * q));
                else v[j][q] = __ldg(p + 32 * q);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int q = 0; q < V; ++q) acc += __uint_as_float(v[j][q].x ^ v[j][q].y ^ v[j][q].z ^ v[j][q].w);
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int ROWB, bool NOALLOC>
void run(const char *name, const uint4 *tab, size_t bytes, float *out, int ctas_per_sm) {
    const uint32_t rows = (uint32_t)(bytes / ROWB);
    const int iters = 64, grid = 148 * ctas_per_sm;
    gather<ROWB, NOALLOC><<<grid, 512>>>(tab, rows, iters, out);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) gather<ROWB, NOALLOC><<<grid, 512>>>(tab, rows, iters, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double moved = 5.0 * grid * 16.0 * iters * 8.0 * ROWB;
    printf("%-28s table %8.1f MB  row %4d B  ctas/SM %d : %8.1f GB/s\n", name, bytes / 1e6, ROWB, ctas_per_sm, moved / ms * 1e-6);
}

int main() {
    uint4 *tab; float *out;
    const size_t big = (size_t)6 << 30;
    cudaMalloc(&tab, big); cudaMalloc(&out, 4);
    cudaMemset(tab, 1, big);
    for (int c = 1; c <= 4; c *= 2) {
        run<512, true>("L2-resident no_allocate", tab, (size_t)48 << 20, out, c);
        run<512, false>("L2-resident ldg", tab, (size_t)48 << 20, out, c);
        run<1536, true>("L2-resident no_allocate", tab, (size_t)48 << 20, out, c);
        run<512, true>("DRAM no_allocate", tab, big, out, c);
        run<1536, true>("DRAM no_allocate", tab, big, out, c);
        run<1536, false>("DRAM ldg", tab, big, out, c);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
