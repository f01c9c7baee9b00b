// Microbenchmarks that size the edge-kernel design (round 1): FFMA vs packed fma.rn.f32x2 vs mma.sync tf32 issue rates
// on sm_100a, plus LDS.128 broadcast.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_ffma(float *out, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    float w = a + threadIdx.x, g = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(g, w, acc[i]);
        w += 1e-9f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float *out, float a, float b) {
    unsigned long long acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = ((unsigned long long)__float_as_uint(threadIdx.x * 0.001f + i) << 32) | __float_as_uint(1.0f * i);
    unsigned long long w = ((unsigned long long)__float_as_uint(a + threadIdx.x) << 32) | __float_as_uint(a);
    unsigned long long g = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(g), "l"(w));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s >> 32)) + __uint_as_float((unsigned)s);
}

__global__ void k_mma_tf32(float *out, float a) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned A[4], B[2];
    for (int j = 0; j < 4; ++j) A[j] = __float_as_uint(a + threadIdx.x + j);
    for (int j = 0; j < 2; ++j) B[j] = __float_as_uint(a * 0.5f + j);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B[0]), "r"(B[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lds128(float *out, int stride) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = (threadIdx.x * stride) & 1023;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 v = sm[(idx + i * 32) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx = (idx + 1) & 1023;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    int nsm = 148, tpb = 512, bps = 2;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, sizeof(float) * nsm * bps * tpb * 4);
    int grid = nsm * bps;
    double warps_per_sm = bps * tpb / 32.0;
    float ms;
    ms = timeit([&] { k_ffma<<<grid, tpb>>>(out, 1.f, 2.f); });
    printf("FFMA    : %.3f ms  %.1f lane-FMA/clk/SM (at %d MHz nominal)\n", ms, warps_per_sm * 32.0 * 16 * ITERS / (ms * 1e-3 * clk_khz * 1e3), clk_khz / 1000);
    ms = timeit([&] { k_ffma2<<<grid, tpb>>>(out, 1.f, 2.f); });
    printf("FFMA2   : %.3f ms  %.1f lane-FMA/clk/SM\n", ms, warps_per_sm * 32.0 * 16 * 2 * ITERS / (ms * 1e-3 * clk_khz * 1e3));
    ms = timeit([&] { k_mma_tf32<<<grid, tpb>>>(out, 1.f); });
    printf("MMA tf32 m16n8k8: %.3f ms  %.1f MAC/clk/SM  (%.1f TFLOP/s chip)\n", ms, warps_per_sm * 1024.0 * 8 * ITERS / (ms * 1e-3 * clk_khz * 1e3),
           2.0 * nsm * warps_per_sm * 1024.0 * 8 * ITERS / (ms * 1e-3) / 1e12);
    for (int stride : {0, 1, 4}) {
        ms = timeit([&] { k_lds128<<<grid, tpb>>>(out, stride); });
        printf("LDS.128 stride %d: %.3f ms  %.1f B/clk/SM\n", stride, ms, warps_per_sm * 32.0 * 16 * 8 * ITERS / (ms * 1e-3 * clk_khz * 1e3));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
