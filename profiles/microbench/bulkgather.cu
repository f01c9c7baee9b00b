// Gather of random ROWB-byte rows into a shared-memory ring with cp.async.bulk (one bulk copy per row and lane), consumed by
// "epilogue" warps that read every row once with LDS: the ceiling of a ring-staged gather (edge kernels, csrc/hn_edge_tc.cu).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulkgather bulkgather.cu && ./bulkgather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// NG groups; group g: 1 producer warp + 4 consumer warps; ring of S slots x B rows of ROWB bytes
template <int ROWB, int NG, int S, int B>
__global__ void __launch_bounds__(32 * NG * 5, 1) ring(const uint8_t *__restrict__ tab, uint32_t rows, int batches, float *out) {
    extern __shared__ __align__(128) uint8_t sm[];
    const uint32_t base = smem_u32(sm);
    const uint32_t bars = base + NG * S * B * ROWB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NG * S; ++i) { mbar_init(bars + 16 * i, 1); mbar_init(bars + 16 * i + 8, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp < NG) {
        const int g = warp;
        uint32_t s = (blockIdx.x * NG + g) * 2654435761u + 777u + lane * 97u;
        for (int it = 0; it < batches; ++it) {
            const int sl = it % S;
            const uint32_t full = bars + 16 * (g * S + sl), empty = full + 8;
            mbar_wait(empty, ((it / S) & 1) ^ 1);
            if (lane == 0) mbar_expect_tx(full, B * ROWB);
            __syncwarp();
            if (lane < B) {
                s = s * 1664525u + 1013904223u;
                const uint32_t r = (uint32_t)(((uint64_t)s * rows) >> 32);
                bulk_g2s(base + ((g * S + sl) * B + lane) * ROWB, tab + (size_t)r * ROWB, ROWB, full);
            }
        }
    } else {
        const int g = (warp - NG) >> 2, q = (warp - NG) & 3;
        float acc = 0.f;
        for (int it = 0; it < batches; ++it) {
            const int sl = it % S;
            const uint32_t full = bars + 16 * (g * S + sl), empty = full + 8;
            mbar_wait(full, (it / S) & 1);
            const float *p = reinterpret_cast<const float *>(sm + ((g * S + sl) * B) * ROWB) + q * 32 + lane;
#pragma unroll
            for (int j = 0; j < B; ++j)
#pragma unroll
                for (int z = 0; z < ROWB / 512; ++z) acc += p[j * (ROWB / 4) + z * 128];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty);
        }
        if (acc == 1.2345f) out[0] = acc;
    }
}

template <int ROWB, int NG, int S, int B>
void run(const uint8_t *tab, size_t bytes, float *out) {
    const uint32_t rows = (uint32_t)(bytes / ROWB);
    const int batches = 4000, smem = NG * S * B * ROWB + 16 * NG * S + 64;
    cudaFuncSetAttribute(ring<ROWB, NG, S, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    ring<ROWB, NG, S, B><<<148, 32 * NG * 5, smem>>>(tab, rows, batches, out);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) ring<ROWB, NG, S, B><<<148, 32 * NG * 5, smem>>>(tab, rows, batches, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double moved = 3.0 * 148 * NG * (double)batches * B * ROWB;
    printf("table %7.1f MB row %4d B  groups %d slots %d x %d rows (ring %3d KB): %8.1f GB/s  %6.1f Mrows/s/SM  %s\n", bytes / 1e6, ROWB, NG, S, B,
           NG * S * B * ROWB / 1024, moved / ms * 1e-6, moved / ROWB / ms * 1e-3 / 148, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint8_t *tab; float *out;
    const size_t big = (size_t)6 << 30;
    cudaMalloc(&tab, big); cudaMalloc(&out, 4);
    cudaMemset(tab, 1, big);
    for (size_t bytes : {(size_t)48 << 20, big}) {
        run<1536, 3, 2, 8>(tab, bytes, out);
        run<1536, 3, 4, 4>(tab, bytes, out);
        run<1536, 3, 2, 4>(tab, bytes, out);
        run<3072, 3, 2, 4>(tab, bytes, out);
        run<3072, 3, 3, 2>(tab, bytes, out);
        run<2560, 3, 2, 4>(tab, bytes, out);
        run<512, 3, 4, 8>(tab, bytes, out);
        run<1536, 2, 4, 4>(tab, bytes, out);
        run<1536, 4, 2, 4>(tab, bytes, out);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
