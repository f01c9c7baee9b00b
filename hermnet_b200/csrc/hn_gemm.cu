// Split-precision dense GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   C[M,N] = A[M,K] . W[N,K]^T (+ bias[N])          fp32 in, fp32 out, fp32-class accuracy
//
// Replaces the cuBLAS fp32 SIMT GEMMs behind the node-side nn.Linear layers of the reference
// (HermNet/rmnet.py:40-49,84-89: x_proj, vec_proj, xvec_proj; hermnet.py:112-116 out_energy) on the fused
// (inference) path.  fp32 parity (BASELINE: 1e-5 relative energies) rules out single-pass TF32, so every operand
// is split  a = hi + lo  with  hi = a & 0xFFFFE000 (exactly representable in TF32)  and  lo = a - hi  (exact in
// fp32, <= 13 significant bits), and the product is accumulated in TMEM as
//     hi.hi + hi.lo + lo.hi                      (3 x tcgen05.mma.kind::tf32, fp32 accumulate)
// the dropped lo.lo term is <= 2^-22 relative.  W is split once on the host side; A is split in shared memory by the
// CTA that consumes it (an elementwise op, so it is oblivious to the 128-byte swizzle TMA wrote).
//
// The kernel is a persistent, warp-specialised CTA per SM (roles and tile shape: see gemm_tf32x3_kernel below).
#include <cuda.h>

#include "hn_common.cuh"

namespace {

constexpr int BM = 128, BK = 32;

// measurement switches for A/B builds (profiles/scripts/gemm_ab.sh); 0 = the product
#ifndef HN_GEMM_AB
#define HN_GEMM_AB 0      // 1: no operand split, 2: no epilogue math / stores, 3: no MMAs
#endif
#ifndef HN_GEMM_MT
#define HN_GEMM_MT 2      // 128-row sub-tiles per CTA tile: they share every W chunk staged in shared memory
#endif
#ifndef HN_GEMM_STAGES
#define HN_GEMM_STAGES (HN_GEMM_MT == 1 ? 3 : 2)
#endif
constexpr int MT = HN_GEMM_MT;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart (SBO = 64), LBO = 1, descriptor version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ScaledSiLU (rmnet.py:110-117) and its derivative in the GEMM epilogues.  The four epilogue warps evaluate 16 K of these
// per 128 x 128 tile; with expf + an IEEE division (~35 instructions each, one warp per scheduler) the epilogue, not the
// tensor pipe, set the tile time of the activation GEMMs (ncu: tensor pipe 14 % active).  ex2.approx / rcp.approx keep
// the relative error at ~3e-7, two orders below the parity tolerance.
__device__ __forceinline__ float sigmoid_fast(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
__device__ __forceinline__ float ssilu(float z) { return z * sigmoid_fast(z) * (1.f / 0.6f); }
__device__ __forceinline__ float dssilu(float z) {
    const float sg = sigmoid_fast(z);
    return sg * (1.f + z * (1.f - sg)) * (1.f / 0.6f);
}

template <int BN>
struct Smem {
    // three operand stages in flight (with two, the tensor pipe sat at 36 %: every K-chunk waited a full TMA round trip)
    static constexpr int kStages = HN_GEMM_STAGES;
    static constexpr int kA = BM * BK * 4;   // 16 KB
    static constexpr int kB = BN * BK * 4;
    static constexpr int kStage = MT * 2 * kA + 2 * kB;      // per sub-tile A raw -> hi | A lo, then W hi | W lo
    static constexpr int kStaging = kStages * kStage;          // 2 x 16 KB: one staging tile per epilogue group
    static constexpr int kBars = kStaging + 2 * 16384;
    static constexpr int kTotal = kBars + 256 + 1024;   // barriers + tmem slot, + slack for 1024-byte alignment
};

// Persistent CTA (one per SM), 512 threads, tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (n-tile fastest).  A tile is
// MT = 2 sub-tiles of 128 rows x BN columns that share every W chunk staged in shared memory: W (hi + lo) is re-read from L2 for
// every tile, twice the bytes of the A chunk, and L2 -> SM bandwidth bounded the main loop (ncu: 8.9 TB/s of it at 60 % tensor
// pipe); with two sub-tiles per W chunk the operand traffic per output row drops by a third.
//   warp 0    : TMA producer (A raw, W_hi, W_lo K-chunks of 32 floats = one 128-byte swizzle row) into a stage ring
//   warp 1    : MMA issuer (one thread); accumulators double-buffered in TMEM (2 x MT x BN columns)
//   warp 2    : TMEM allocation
//   warps 4-7 : split A in place (hi) + side buffer (lo)
//   warps 8-15: epilogue of the PREVIOUS tile while the next one is loaded and multiplied (two groups of four warps, each
//               with its own staging tile): tcgen05.ld -> +bias / activation -> swizzled staging tile in smem -> TMA store
template <int BN>
__global__ void __launch_bounds__(512, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                   const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmC,
                   const __grid_constant__ CUtensorMap tmC2, const float *__restrict__ bias, float *__restrict__ C,
                   int M, int N, int K, long long ldc, int mode, const float *__restrict__ aux, long long ld_aux,
                   float *__restrict__ C2, long long ldc2) {
    extern __shared__ uint8_t smem_raw[];
    using L = Smem<BN>;
    constexpr int STAGES = L::kStages;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + L::kBars;
    const uint32_t full0 = bars, split0 = bars + 8 * STAGES, empty0 = bars + 16 * STAGES, tfull0 = bars + 24 * STAGES,
                   tempty0 = tfull0 + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen + L::kBars + 24 * STAGES + 40);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = K / BK;
    const int n_tiles_n = N / BN;
    const long long n_tiles = (long long)n_tiles_n * ((M + BM * MT - 1) / (BM * MT));

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(split0 + 8 * s, 128);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        constexpr uint32_t cols = 2 * MT * BN;   // power of two >= 32 (BN in {64, 128}, MT in {1, 2})
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            long long kbg = 0;                                  // running K-chunk counter (stage ring position)
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int n0 = (int)(tile % n_tiles_n) * BN, m0 = (int)(tile / n_tiles_n) * (BM * MT);
                for (int kb = 0; kb < num_k; ++kb, ++kbg) {
                    const int s = (int)(kbg % STAGES);
                    const uint32_t ph = (uint32_t)(kbg / STAGES) & 1;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    const uint32_t st = base + s * L::kStage;
                    mbar_expect_tx(full0 + 8 * s, MT * L::kA + 2 * L::kB);
                    for (int h = 0; h < MT; ++h)      // (rows beyond M: zero-filled by the TMA unit, full box counted)
                        tma_load_2d(st + h * 2 * L::kA, &tmA, full0 + 8 * s, kb * BK, m0 + h * BM);
                    tma_load_2d(st + MT * 2 * L::kA, &tmBhi, full0 + 8 * s, kb * BK, n0);
                    tma_load_2d(st + MT * 2 * L::kA + L::kB, &tmBlo, full0 + 8 * s, kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // instruction descriptor: D=F32 (bit 4), A=B=TF32 (2<<7, 2<<10), K-major A and B, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        long long kbg = 0, it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int a = (int)(it & 1);
            mbar_wait(tempty0 + 8 * a, (uint32_t)((it >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + (uint32_t)(a * MT * BN);
            for (int kb = 0; kb < num_k; ++kb, ++kbg) {
                const int s = (int)(kbg % STAGES);
                const uint32_t ph = (uint32_t)(kbg / STAGES) & 1;
                mbar_wait(full0 + 8 * s, ph);
                mbar_wait(split0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t st = base + s * L::kStage;
                    const uint32_t b_hi = st + MT * 2 * L::kA, b_lo = b_hi + L::kB;
#pragma unroll
                    for (int h = 0; h < MT; ++h) {
                        const uint32_t a_hi = st + h * 2 * L::kA, a_lo = a_hi + L::kA, th = tacc + (uint32_t)(h * BN);
#pragma unroll
                        for (int k4 = 0; k4 < (HN_GEMM_AB == 3 ? 0 : BK / 8); ++k4) {
                            const uint32_t off = k4 * 32;   // 8 tf32 = 32 bytes inside the 128-byte swizzle row
                            umma_tf32(th, umma_desc(a_hi + off), umma_desc(b_hi + off), idesc, (kb | k4) != 0);
                            umma_tf32(th, umma_desc(a_hi + off), umma_desc(b_lo + off), idesc, 1);
                            umma_tf32(th, umma_desc(a_lo + off), umma_desc(b_hi + off), idesc, 1);
                        }
                    }
                    umma_commit(empty0 + 8 * s);                          // stage free once these MMAs retire
                    if (kb == num_k - 1) umma_commit(tfull0 + 8 * a);     // accumulator complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 8) {
        const int t = threadIdx.x - 128;
        long long kbg = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < num_k; ++kb, ++kbg) {
                const int s = (int)(kbg % STAGES);
                const uint32_t ph = (uint32_t)(kbg / STAGES) & 1;
                mbar_wait(full0 + 8 * s, ph);
#pragma unroll
                for (int sub = 0; sub < MT; ++sub) {
                float4 *hi = reinterpret_cast<float4 *>(gen + s * L::kStage + sub * 2 * L::kA);
                float4 *lo = reinterpret_cast<float4 *>(gen + s * L::kStage + sub * 2 * L::kA + L::kA);
#pragma unroll
                for (int i = 0; i < (HN_GEMM_AB == 1 ? 0 : L::kA / 16 / 128); ++i) {
                    const float4 v = hi[t + i * 128];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    hi[t + i * 128] = h;
                    lo[t + i * 128] = l;
                }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
                mbar_arrive(split0 + 8 * s);
            }
        }
    } else if (warp >= 8) {
        // epilogue: two groups of four warps; warp w owns TMEM lanes 32*(w%4) .. +31 = output rows m0 + 32*(w%4) + lane, group
        // g = (w-8)/4 takes the 32-column chunks g, g+2, ... of the tile (one chunk: tcgen05.ld -> bias / activation -> a
        // 16 KB staging tile in the 128-byte-swizzle layout -> ONE TMA store: full 128-byte lines, rows beyond M clipped).
        // Two groups because a chunk is a serial chain (TMEM load, the pre-activation load of mode 2, MUFU, the staging
        // barriers, the wait for the previous store to have left the staging tile): with one group the epilogue, not the
        // tensor pipe or HBM, set the tile time of every GEMM with K <= 384 (profiles/r2: A/B without epilogue 5.2 -> 3.0 ms).
        const int t = (threadIdx.x - 256) & 127;
        const int grp = (threadIdx.x - 256) >> 7;
        const int q = warp & 3;
        const int rl = q * 32 + lane;                     // row inside the tile
        const bool two = (mode == 1 && C2 != nullptr);
        uint8_t *sC = gen + L::kStaging + grp * 16384;    // one staging tile per group
        const int bar_id = 1 + grp;
        long long it = 0;
        constexpr int CH = BN / 64;                          // chunks of a sub-tile per group
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int n0 = (int)(tile % n_tiles_n) * BN, m0 = (int)(tile / n_tiles_n) * (BM * MT);
            const int n_sub = min(MT, (M - m0 + BM - 1) / BM);       // sub-tiles with rows inside the matrix
            const int a = (int)(it & 1);
            float4 z[8];
            if (mode == 2 && (long long)m0 + rl < M) {      // pre-activations of the group's first chunk: in flight while the tile is multiplied
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    z[j] = __ldg(reinterpret_cast<const float4 *>(aux + ((long long)m0 + rl) * ld_aux + n0 + grp * 32 + 4 * j));
            }
            mbar_wait(tfull0 + 8 * a, (uint32_t)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int ci = 0; ci < n_sub * CH; ++ci) {
                const int h = ci / CH, c0 = grp * 32 + (ci % CH) * 64;
                const int mh = m0 + h * BM;
                const long long row = (long long)mh + rl;
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * MT * BN + h * BN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ci + 1 == n_sub * CH) {       // this warp has read its share of the accumulators: hand them back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + 8 * a);            // (8 arrivals = all epilogue warps)
                }
                if (HN_GEMM_AB == 2) continue;
                float4 o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                       __uint_as_float(r[4 * j + 3]));
                    if (bias != nullptr) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + n0 + c0 + 4 * j));
                        o[j].x += b.x; o[j].y += b.y; o[j].z += b.z; o[j].w += b.w;
                    }
                }
                // the TMA store that read the staging tile last must be done reading it
                if (t == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(0) : "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (two) {                // C2 = pre-activation first, through the same staging tile
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4 *>(sC + rl * 128 + ((j ^ (rl & 7)) << 4)) = o[j];   // 128-byte swizzle: chunk ^= row % 8
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (t == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC2),
                                     "r"(smem_u32(sC)), "r"(n0 + c0), "r"(mh)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (mode == 1) {          // C = ScaledSiLU(pre)   (rmnet.py:110-117)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        o[j].x = ssilu(o[j].x); o[j].y = ssilu(o[j].y); o[j].z = ssilu(o[j].z); o[j].w = ssilu(o[j].w);
                    }
                } else if (mode == 2) {   // C = acc * ScaledSiLU'(aux): backward through the activation
                    if (row < M) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            o[j].x *= dssilu(z[j].x); o[j].y *= dssilu(z[j].y); o[j].z *= dssilu(z[j].z); o[j].w *= dssilu(z[j].w);
                        }
                    }
                    if (ci + 1 < n_sub * CH) {     // next chunk's pre-activations
                        const int hn = (ci + 1) / CH, cn = grp * 32 + ((ci + 1) % CH) * 64;
                        const long long rn = (long long)m0 + hn * BM + rl;
                        if (rn < M) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                z[j] = __ldg(reinterpret_cast<const float4 *>(aux + rn * ld_aux + n0 + cn + 4 * j));
                        }
                    }
                }
                if (two) {
                    if (t == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(0) : "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(sC + rl * 128 + ((j ^ (rl & 7)) << 4)) = o[j];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> TMA reads
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (t == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmC),
                                 "r"(smem_u32(sC)), "r"(n0 + c0), "r"(mh)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (t == 0) asm volatile("cp.async.bulk.wait_group %0;" ::"n"(0) : "memory");   // writes complete before exit
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        constexpr uint32_t cols = 2 * MT * BN;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, K] fp32 row-major with row pitch ld (floats): box = 32 floats (128 B, one swizzle row) x box_rows
int make_map(CUtensorMap *map, const float *ptr, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2;
}

template <int BN>
int launch(const float *A, int64_t M, int64_t K, int64_t lda, const float *Whi, const float *Wlo, int64_t N, const float *bias,
           float *C, int64_t ldc, int mode, const float *aux, int64_t ld_aux, float *C2, int64_t ldc2, cudaStream_t st) {
    const char *where = "hn_gemm_tf32x3";
    CUtensorMap ta, tbh, tbl;
    HN_REQUIRE(make_map(&ta, A, M, K, lda, BM) == 0, where, "cuTensorMapEncodeTiled failed for A");
    HN_REQUIRE(make_map(&tbh, Whi, N, K, K, BN) == 0, where, "cuTensorMapEncodeTiled failed for W_hi");
    HN_REQUIRE(make_map(&tbl, Wlo, N, K, K, BN) == 0, where, "cuTensorMapEncodeTiled failed for W_lo");
    CUtensorMap tc, tc2;
    HN_REQUIRE(make_map(&tc, C, M, N, ldc, BM) == 0, where, "cuTensorMapEncodeTiled failed for C");
    if (C2 != nullptr) HN_REQUIRE(make_map(&tc2, C2, M, N, ldc2, BM) == 0, where, "cuTensorMapEncodeTiled failed for C2");
    else tc2 = tc;
    // (a per-device attribute: set on every launch instead of remembering it in library state)
    HN_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::kTotal), where);
    const long long tiles = (N / BN) * ((M + BM * MT - 1) / (BM * MT));
    const int sms = hn::num_sms() > 0 ? hn::num_sms() : 148;
    dim3 grid((unsigned)(tiles < sms ? tiles : sms));
    gemm_tf32x3_kernel<BN><<<grid, 512, Smem<BN>::kTotal, st>>>(ta, tbh, tbl, tc, tc2, bias, C, (int)M, (int)N, (int)K, (long long)ldc, mode, aux,
                                                                 (long long)ld_aux, C2, (long long)ldc2);
    return hn::check_launch(where);
}

}  // namespace

// C[M,N] (row pitch ldc) = epilogue(A[M,K] (row pitch lda) . W[N,K]^T + bias);  W given pre-split (hi = W & 0xFFFFE000, lo = W - hi).
//   mode 0: identity            mode 1: C = ScaledSiLU(pre), C2 = pre (optional, row pitch ldc2)
//   mode 2: C = (A.W^T) * ScaledSiLU'(aux[row][col])  (aux row pitch ld_aux) -- the backward pass through the activation
// Requirements: K % 32 == 0, N % 64 == 0, row pitches % 4 == 0, 16-byte aligned pointers.
extern "C" int hn_gemm_tf32x3_ex(const float *A, int64_t M, int64_t K, int64_t lda, const float *W_hi, const float *W_lo, int64_t N,
                                 const float *bias, float *C, int64_t ldc, int32_t mode, const float *aux, int64_t ld_aux,
                                 float *C2, int64_t ldc2, void *stream) {
    const char *where = "hn_gemm_tf32x3";
    if (M <= 0) return 0;
    HN_REQUIRE(K >= 32 && K % 32 == 0, where, "K must be a positive multiple of 32");
    HN_REQUIRE(N >= 64 && N % 64 == 0, where, "N must be a positive multiple of 64");
    HN_REQUIRE(lda % 4 == 0 && ldc % 4 == 0 && ld_aux % 4 == 0 && ldc2 % 4 == 0, where, "row pitches must be multiples of 4 floats");
    HN_REQUIRE(M < (1ll << 31), where, "M out of range");
    HN_REQUIRE(mode >= 0 && mode <= 2, where, "unknown epilogue mode");
    HN_REQUIRE(mode != 2 || aux != nullptr, where, "mode 2 needs the pre-activation (aux)");
    HN_REQUIRE((((uintptr_t)A | (uintptr_t)W_hi | (uintptr_t)W_lo | (uintptr_t)C | (uintptr_t)bias | (uintptr_t)aux | (uintptr_t)C2) & 15) == 0,
               where, "pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (N % 128 == 0) return launch<128>(A, M, K, lda, W_hi, W_lo, N, bias, C, ldc, mode, aux, ld_aux, C2, ldc2, st);
    return launch<64>(A, M, K, lda, W_hi, W_lo, N, bias, C, ldc, mode, aux, ld_aux, C2, ldc2, st);
}

extern "C" int hn_gemm_tf32x3(const float *A, int64_t M, int64_t K, int64_t lda, const float *W_hi, const float *W_lo, int64_t N,
                              const float *bias, float *C, int64_t ldc, void *stream) {
    return hn_gemm_tf32x3_ex(A, M, K, lda, W_hi, W_lo, N, bias, C, ldc, 0, nullptr, 4, nullptr, 4, stream);
}
