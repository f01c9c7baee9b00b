// Row-group PaiNN edge kernels with a piecewise-polynomial radial filter: forward and destination-major backward.
//
// Same arithmetic as hn_edge.cu (reference rmnet.py:55-73 + the Gaussian RBF x polynomial envelope of rmnet.py:168-193):
//   phi_e = W[m] . (env(u) * gauss_k(u)) + b[m],   u = d_e / rc
// The sum over basis functions  f_c(u) = sum_k W[m][k][c] exp(coeff (u - offset_k)^2)  is a smooth function of the
// ONE scalar u.  Between two grid points offset_kc <= u < offset_kc+1 it is tabulated as a degree-9 polynomial in the
// local coordinate s = 2 (u - offset_kc)(K-1) - 1  (table built in fp64 from ALL K basis functions by
// hermnet_b200/filter_table.py; its fp32 Horner evaluation is closer to the exact sum than the reference's own fp32
// evaluation of the K exponentials: 8e-8 vs 2e-7 relative, derivative 1.8e-7 vs 2.8e-7).  Evaluating phi then costs
// 9 FMAs per channel instead of 12 (band) / 128 (reference) and needs no exponentials.
//
// The row-per-warp kernels are bound by the L1/LSU pipe: every edge re-reads its filter rows (18 KB at F=128).  Here
// the graph builder groups 8 rows of one sub-network and sorts the group's edges by grid interval kc; one warp owns
// one (group, 64-channel slice), keeps the interval's 10 x 3 coefficient pairs in registers and re-loads them only when
// kc changes (~4 edges share one load).  The 8 rows' accumulators live in registers (selected by a switch on the local
// row) -- no shared memory, no atomics, deterministic order.  The interval is re-derived from the CURRENT distance, so
// a stale plan (graph re-used after the atoms moved) only costs extra coefficient loads, never accuracy.
#include "hn_common.cuh"

namespace {

typedef unsigned long long u64;

constexpr int kGR = 8;      // rows per group
constexpr int kNC = 10;     // polynomial coefficients per interval (degree 9)

__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float hsum(u64 v) {
    float lo, hi;
    upk(v, lo, hi);
    return lo + hi;
}
__device__ __forceinline__ u64 ldg64(const float *p) { return __ldg(reinterpret_cast<const u64 *>(p)); }

// polynomial envelope (rmnet.py:183-193): env(u), d env/du
template <bool DERIV>
__device__ __forceinline__ void envelope(float u, int p, float &env, float &denv) {
    const float a = -0.5f * (float)((p + 1) * (p + 2)), b = (float)(p * (p + 2)), c = -0.5f * (float)(p * (p + 1));
    float um = 1.f;
    if (p == 5) um = (u * u) * (u * u);
    else for (int i = 0; i < p - 1; ++i) um *= u;
    const float u0 = um * u, u1 = u0 * u, u2 = u1 * u;
    env = 1.f + a * u0 + b * u1 + c * u2;
    if (DERIV) denv = a * (float)p * um + b * (float)(p + 1) * u0 + c * (float)(p + 2) * u1;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(unsigned dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Per-warp shared memory: two header batches (32 slots x {meta int4, geom float4}) + a ring of kDepth gathered edges
// (6 x 64-channel slices each).  Everything is filled with cp.async, so the loads of the next kDepth edges are in flight
// while an edge is processed and no registers are spent on prefetching.
constexpr int kDepth = 6;
constexpr int kWarps = 8;       // warps per CTA: consecutive (group, slice) units that sweep the intervals in phase
constexpr int kHdrBytes = 2 * 32 * 32;
constexpr int kEdgeBytes = 6 * 256;
constexpr int kWarpBytes = kHdrBytes + kDepth * kEdgeBytes;      // 11264

#define HN_LOAD_COEF(kc)                                                                              \
    {                                                                                                 \
        const float *c_ = Cm + (size_t)(kc) * (kNC * F3);                                             \
        _Pragma("unroll") for (int n_ = 0; n_ < kNC; ++n_) {                                          \
            cf[n_][0] = ldg64(c_ + n_ * F3);                                                          \
            cf[n_][1] = ldg64(c_ + n_ * F3 + F);                                                      \
            cf[n_][2] = ldg64(c_ + n_ * F3 + 2 * F);                                                  \
        }                                                                                             \
        off_kc = __ldg(offset + (kc));                                                                \
        kc_cur = (kc);                                                                                \
    }

#define HN_SWITCH8(lrow, STMT)                                                                        \
    switch (lrow) {                                                                                   \
        case 0: { constexpr int R_ = 0; STMT } break;                                                 \
        case 1: { constexpr int R_ = 1; STMT } break;                                                 \
        case 2: { constexpr int R_ = 2; STMT } break;                                                 \
        case 3: { constexpr int R_ = 3; STMT } break;                                                 \
        case 4: { constexpr int R_ = 4; STMT } break;                                                 \
        case 5: { constexpr int R_ = 5; STMT } break;                                                 \
        case 6: { constexpr int R_ = 6; STMT } break;                                                 \
        case 7: { constexpr int R_ = 7; STMT } break;                                                 \
        default: break;                                                                               \
    }

// header batch b (slots E0+32b ..) -> hbuf[b & 1]; lane l fetches slot l of the batch
#define HN_ISSUE_HDR(b)                                                                               \
    {                                                                                                 \
        const int s_ = E0 + 32 * (b) + lane;                                                          \
        if (s_ < E1) {                                                                                \
            const unsigned d_ = smem_u32(wsm + ((b) & 1) * 1024 + lane * 32);                         \
            cp_async16(d_, meta + s_);                                                                \
            cp_async16(d_ + 16, geom_g + s_);                                                         \
        }                                                                                             \
    }
// gathered rows of relative edge j -> ring stage j % kDepth (reads the edge's header from shared memory).  The edge
// record is 6 segments of 256 B (xh parts a,b,c and vec x,y,z of this 64-channel slice) = 96 chunks of 16 B; lane l
// copies chunks l, l+32, l+64 with cp.async.cg (L2 only: keeps L1 for the filter table)
#define HN_ISSUE_EDGE(j)                                                                              \
    {                                                                                                 \
        if ((j) < n) {                                                                                \
            const int4 mj_ = *reinterpret_cast<const int4 *>(wsm + (((j) >> 5) & 1) * 1024 + ((j) & 31) * 32); \
            const float *x_ = xh + (size_t)mj_.z * F3 + slice * 64, *v_ = vec + (size_t)mj_.x * F3 + slice * 64; \
            const unsigned d_ = smem_u32(wsm + kHdrBytes + ((j) % kDepth) * kEdgeBytes) + lane * 16;  \
            const int seg_ = lane >> 4, o_ = (lane & 15) * 4;      /* chunk l: segment l/16, floats 4*(l%16) */ \
            cp_async16(d_, x_ + seg_ * F + o_);                          /* segments 0,1: xh parts a,b */   \
            cp_async16(d_ + 512, (seg_ == 0 ? x_ + 2 * F : v_) + o_);    /* segments 2,3: xh part c, vec x */ \
            cp_async16(d_ + 1024, v_ + (seg_ + 1) * F + o_);             /* segments 4,5: vec y,z */        \
        }                                                                                             \
    }

// ---------------------------------------------------------------------------------------------------------------
// forward.  meta[slot] = (source atom, local row 0..7, xh row, interval), geom_g[slot] = (ux, uy, uz, d), slots of
// group g are [gptr[g], gptr[g+1]) sorted by grid interval.
// ---------------------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(32 * kWarps, 1)
edge_fwd_group_kernel(const hn_edge_params P, const float *__restrict__ xh, const float *__restrict__ vec,
                      const float4 *__restrict__ geom_g, const int *__restrict__ gptr, const int4 *__restrict__ meta,
                      const int *__restrict__ group_rows, const int *__restrict__ group_mod, int n_groups,
                      const float *__restrict__ coef, const float *__restrict__ bias, const float *__restrict__ offset,
                      float *__restrict__ dx, float *__restrict__ dvec) {
    extern __shared__ __align__(16) char smem_raw[];
    constexpr int F = 64 * NS, F3 = 3 * F;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kWarps + warp;
    const int group = unit / NS, slice = unit - group * NS;
    if (group >= n_groups) return;
    char *wsm = smem_raw + warp * kWarpBytes;
    const int m = __ldg(group_mod + group);
    const int K = P.num_rbf;
    const int ch = slice * 64 + 2 * lane;
    const float *Cm = coef + (size_t)m * (K - 1) * (kNC * F3) + ch;
    const u64 ba = ldg64(bias + (size_t)m * F3 + ch), bb = ldg64(bias + (size_t)m * F3 + F + ch),
              bc = ldg64(bias + (size_t)m * F3 + 2 * F + ch);
    const float c1 = 1.0f / sqrtf(3.0f * (float)F), c2 = 1.0f / sqrtf((float)F);
    const u64 c1p = pk(c1, c1), c2p = pk(c2, c2);
    const float Km1 = (float)(K - 1);
    const int E0 = __ldg(gptr + group), E1 = __ldg(gptr + group + 1);
    const int n = E1 - E0;
    u64 acc[kGR][4];
#pragma unroll
    for (int r = 0; r < kGR; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0ull;
    u64 cf[kNC][3];
#pragma unroll
    for (int q = 0; q < kNC; ++q) cf[q][0] = cf[q][1] = cf[q][2] = 0ull;
    int kc_cur = -1;
    float off_kc = 0.f;

    if (n > 0) {
        HN_ISSUE_HDR(0)
        cp_commit();
        cp_wait<0>();
        __syncwarp();
        HN_ISSUE_HDR(1)
#pragma unroll 1
        for (int j = 0; j < kDepth; ++j) {
            HN_ISSUE_EDGE(j)
            cp_commit();
        }
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            cp_wait<kDepth - 1>();
            __syncwarp();
            const char *hp = wsm + ((i >> 5) & 1) * 1024 + (i & 31) * 32;
            const int lrow = *reinterpret_cast<const int *>(hp + 4);
            const float4 g = *reinterpret_cast<const float4 *>(hp + 16);
            const u64 *rq = reinterpret_cast<const u64 *>(wsm + kHdrBytes + (i % kDepth) * kEdgeBytes) + lane;
            const u64 R0 = rq[0], R1 = rq[32], R2 = rq[64], R3 = rq[96], R4 = rq[128], R5 = rq[160];
            const float u = g.w * P.inv_rc;
            u64 fa = ba, fb = bb, fc = bc;
            if (u < 1.f) {
                const int kc = min((int)(u * Km1), K - 2);
                if (kc != kc_cur) HN_LOAD_COEF(kc)
                const float s = fmaf(2.f * Km1, u - off_kc, -1.f);
                const u64 sp = pk(s, s);
                u64 pa = cf[kNC - 1][0], pb = cf[kNC - 1][1], pc = cf[kNC - 1][2];
#pragma unroll
                for (int q = kNC - 2; q >= 0; --q) {
                    pa = fma2(pa, sp, cf[q][0]);
                    pb = fma2(pb, sp, cf[q][1]);
                    pc = fma2(pc, sp, cf[q][2]);
                }
                float env, denv;
                envelope<false>(u, P.env_p, env, denv);
                const u64 ep = pk(env, env);
                fa = fma2(ep, pa, ba), fb = fma2(ep, pb, bb), fc = fma2(ep, pc, bc);
            }
            const u64 tb = mul2(mul2(R1, fb), c1p);
            const u64 tc = mul2(mul2(R2, fc), c2p);
            const u64 m0 = mul2(R0, fa);
            const u64 m1 = fma2(tc, pk(g.x, g.x), mul2(R3, tb));
            const u64 m2 = fma2(tc, pk(g.y, g.y), mul2(R4, tb));
            const u64 m3 = fma2(tc, pk(g.z, g.z), mul2(R5, tb));
            HN_SWITCH8(lrow, acc[R_][0] = add2(acc[R_][0], m0); acc[R_][1] = add2(acc[R_][1], m1);
                       acc[R_][2] = add2(acc[R_][2], m2); acc[R_][3] = add2(acc[R_][3], m3);)
            __syncwarp();                                   // stage i % kDepth and (at batch ends) the header buffer are free
            if ((i & 31) == 31) HN_ISSUE_HDR((i >> 5) + 2)  // batch b is done: fetch batch b+2 into its buffer
            HN_ISSUE_EDGE(i + kDepth)
            cp_commit();
        }
        cp_wait<0>();
    }
#pragma unroll
    for (int r = 0; r < kGR; ++r) {
        const int row = __ldg(group_rows + (size_t)group * kGR + r);
        if (row < 0) continue;
        *reinterpret_cast<u64 *>(dx + (size_t)row * F + ch) = acc[r][0];
        *reinterpret_cast<u64 *>(dvec + (size_t)row * F3 + ch) = acc[r][1];
        *reinterpret_cast<u64 *>(dvec + (size_t)row * F3 + F + ch) = acc[r][2];
        *reinterpret_cast<u64 *>(dvec + (size_t)row * F3 + 2 * F + ch) = acc[r][3];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// destination-major backward: per slot (dL/du_x, dL/du_y, dL/du_z, dL/dd), partial over this 64-channel slice, written
// in slot order (the caller adds the slices and un-permutes).
// ---------------------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(32 * kWarps, 1)
edge_bwd_dst_group_kernel(const hn_edge_params P, const float *__restrict__ xh, const float *__restrict__ vec,
                          const float4 *__restrict__ geom_g, const int *__restrict__ gptr, const int4 *__restrict__ meta,
                          const int *__restrict__ group_rows, const int *__restrict__ group_mod, int n_groups,
                          const float *__restrict__ coef, const float *__restrict__ bias, const float *__restrict__ offset,
                          const float *__restrict__ g_dx, const float *__restrict__ g_dvec, float *__restrict__ g_geom_g,
                          long long n_slots) {
    extern __shared__ __align__(16) char smem_raw[];
    constexpr int F = 64 * NS, F3 = 3 * F;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kWarps + warp;
    const int group = unit / NS, slice = unit - group * NS;
    if (group >= n_groups) return;
    char *wsm = smem_raw + warp * kWarpBytes;
    const int m = __ldg(group_mod + group);
    const int K = P.num_rbf;
    const int ch = slice * 64 + 2 * lane;
    const float *Cm = coef + (size_t)m * (K - 1) * (kNC * F3) + ch;
    const u64 bc = ldg64(bias + (size_t)m * F3 + 2 * F + ch);
    const float c1 = 1.0f / sqrtf(3.0f * (float)F), c2 = 1.0f / sqrtf((float)F);
    const u64 c1p = pk(c1, c1), c2p = pk(c2, c2);
    const float Km1 = (float)(K - 1);
    const int E0 = __ldg(gptr + group), E1 = __ldg(gptr + group + 1);
    const int n = E1 - E0;
    if (n <= 0) return;
    float *out = g_geom_g + ((size_t)slice * n_slots + E0) * 4;
    u64 gr[kGR][4];      // upstream gradients of the group's rows: g_dx, g_dvec x/y/z
#pragma unroll
    for (int r = 0; r < kGR; ++r) {
        const int row = __ldg(group_rows + (size_t)group * kGR + r);
        if (row >= 0) {
            gr[r][0] = ldg64(g_dx + (size_t)row * F + ch);
            gr[r][1] = ldg64(g_dvec + (size_t)row * F3 + ch);
            gr[r][2] = ldg64(g_dvec + (size_t)row * F3 + F + ch);
            gr[r][3] = ldg64(g_dvec + (size_t)row * F3 + 2 * F + ch);
        } else {
            gr[r][0] = gr[r][1] = gr[r][2] = gr[r][3] = 0ull;
        }
    }
    u64 cf[kNC][3];
#pragma unroll
    for (int q = 0; q < kNC; ++q) cf[q][0] = cf[q][1] = cf[q][2] = 0ull;
    int kc_cur = -1;
    float off_kc = 0.f;

    HN_ISSUE_HDR(0)
    cp_commit();
    cp_wait<0>();
    __syncwarp();
    HN_ISSUE_HDR(1)
#pragma unroll 1
    for (int j = 0; j < kDepth; ++j) {
        HN_ISSUE_EDGE(j)
        cp_commit();
    }
    float r8[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) r8[q] = 0.f;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        cp_wait<kDepth - 1>();
        __syncwarp();
        const char *hp = wsm + ((i >> 5) & 1) * 1024 + (i & 31) * 32;
        const int lrow = *reinterpret_cast<const int *>(hp + 4);
        const float4 g = *reinterpret_cast<const float4 *>(hp + 16);
        const u64 *rq = reinterpret_cast<const u64 *>(wsm + kHdrBytes + (i % kDepth) * kEdgeBytes) + lane;
        const u64 R0 = rq[0], R1 = rq[32], R2 = rq[64], R3 = rq[96], R4 = rq[128], R5 = rq[160];
        const float u = g.w * P.inv_rc;
        u64 fc = bc, da = 0ull, db = 0ull, dc = 0ull;
        if (u < 1.f) {
            const int kc = min((int)(u * Km1), K - 2);
            if (kc != kc_cur) HN_LOAD_COEF(kc)
            const float s = fmaf(2.f * Km1, u - off_kc, -1.f);
            const u64 sp = pk(s, s);
            u64 pa = cf[kNC - 1][0], pb = cf[kNC - 1][1], pc = cf[kNC - 1][2];
            u64 qa = 0ull, qb = 0ull, qc = 0ull;        // d poly / ds
#pragma unroll
            for (int q = kNC - 2; q >= 0; --q) {
                qa = fma2(qa, sp, pa);
                qb = fma2(qb, sp, pb);
                qc = fma2(qc, sp, pc);
                pa = fma2(pa, sp, cf[q][0]);
                pb = fma2(pb, sp, cf[q][1]);
                pc = fma2(pc, sp, cf[q][2]);
            }
            float env, denv;
            envelope<true>(u, P.env_p, env, denv);
            const u64 ep = pk(env, env);
            const float k1 = denv * P.inv_rc, k2 = env * 2.f * Km1 * P.inv_rc;
            const u64 k1p = pk(k1, k1), k2p = pk(k2, k2);
            fc = fma2(ep, pc, bc);
            da = fma2(k2p, qa, mul2(k1p, pa));
            db = fma2(k2p, qb, mul2(k1p, pb));
            dc = fma2(k2p, qc, mul2(k1p, pc));
        }
        u64 gx = 0ull, gv0 = 0ull, gv1 = 0ull, gv2 = 0ull;
        HN_SWITCH8(lrow, gx = gr[R_][0]; gv0 = gr[R_][1]; gv1 = gr[R_][2]; gv2 = gr[R_][3];)
        const u64 tb = mul2(fma2(gv2, R5, fma2(gv1, R4, mul2(gv0, R3))), c1p);                          // dL/d(Pb*phib)
        const u64 tc = mul2(fma2(gv2, pk(g.z, g.z), fma2(gv1, pk(g.y, g.y), mul2(gv0, pk(g.x, g.x)))), c2p);
        const u64 gd = fma2(mul2(tc, R2), dc, fma2(mul2(tb, R1), db, mul2(mul2(gx, R0), da)));
        const u64 cphi = mul2(mul2(R2, fc), c2p);
        const float v0 = hsum(mul2(gv0, cphi)), v1 = hsum(mul2(gv1, cphi)), v2 = hsum(mul2(gv2, cphi)), v3 = hsum(gd);
        if ((i & 1) == 0) { r8[0] = v0; r8[1] = v1; r8[2] = v2; r8[3] = v3; }
        else { r8[4] = v0; r8[5] = v1; r8[6] = v2; r8[7] = v3; }
        if ((i & 1) || i == n - 1) {
            // 8 values x 32 lanes -> 8 sums with 9 shuffles (halving butterfly); slots i-1 (r8[0..3]) and i (r8[4..7])
            const bool two = (i & 1);
            if (!two) { r8[4] = r8[5] = r8[6] = r8[7] = 0.f; }
            float a4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float send = b4 ? r8[q] : r8[4 + q];
                const float keep = b4 ? r8[4 + q] : r8[q];
                a4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
            float a2[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float send = b3 ? a4[q] : a4[2 + q];
                const float keep = b3 ? a4[2 + q] : a4[q];
                a2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            float a1;
            {
                const float send = b2 ? a2[0] : a2[1];
                const float keep = b2 ? a2[1] : a2[0];
                a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
            const int e = two ? i - 1 : i;
            if ((lane & 3) == 0 && (two || !b4)) out[(size_t)(e + (b4 ? 1 : 0)) * 4 + (b3 ? 2 : 0) + (b2 ? 1 : 0)] = a1;
        }
        __syncwarp();
        if ((i & 31) == 31) HN_ISSUE_HDR((i >> 5) + 2)
        HN_ISSUE_EDGE(i + kDepth)
        cp_commit();
    }
    cp_wait<0>();
}

int validate_group(const char *where, const hn_edge_params *p) {
    HN_REQUIRE(p != nullptr, where, "null params");
    HN_REQUIRE(p->hidden % 64 == 0 && p->hidden >= 64 && p->hidden <= 512, where,
               "row-group edge kernels need hidden_channels in {64,128,...,512}");
    HN_REQUIRE(p->num_rbf >= 2, where, "num_rbf must be >= 2");
    HN_REQUIRE(p->env_p >= 1, where, "envelope exponent must be >= 1");
    return 0;
}

}  // namespace

#define HN_DISPATCH_NS(F, ...)                                       \
    switch ((F) / 64) {                                              \
        case 1: { constexpr int NS = 1; __VA_ARGS__; break; }        \
        case 2: { constexpr int NS = 2; __VA_ARGS__; break; }        \
        case 3: { constexpr int NS = 3; __VA_ARGS__; break; }        \
        case 4: { constexpr int NS = 4; __VA_ARGS__; break; }        \
        case 5: { constexpr int NS = 5; __VA_ARGS__; break; }        \
        case 6: { constexpr int NS = 6; __VA_ARGS__; break; }        \
        case 7: { constexpr int NS = 7; __VA_ARGS__; break; }        \
        default: { constexpr int NS = 8; __VA_ARGS__; break; }       \
    }

extern "C" int32_t hn_painn_edge_group_supported(int32_t hidden, int32_t num_rbf) {
    return (hidden % 64 == 0 && hidden >= 64 && hidden <= 512 && num_rbf >= 2) ? 1 : 0;
}

extern "C" int hn_painn_edge_fwd_group(const hn_edge_params *p, const float *xh, const float *vec, const float *geom_g,
                                       const int32_t *gptr, const int32_t *meta, const int32_t *group_rows,
                                       const int32_t *group_mod, int32_t n_groups, const float *coef, const float *bias,
                                       const float *offset, float *dx, float *dvec, void *stream) {
    const char *where = "hn_painn_edge_fwd_group";
    if (int rc = validate_group(where, p)) return rc;
    if (n_groups <= 0) return 0;
    const long long units = (long long)n_groups * (p->hidden / 64);
    dim3 grid((unsigned)((units + kWarps - 1) / kWarps));
    HN_DISPATCH_NS(p->hidden, HN_CUDA(cudaFuncSetAttribute(edge_fwd_group_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           kWarps * kWarpBytes), where);
                   edge_fwd_group_kernel<NS><<<grid, 32 * kWarps, kWarps * kWarpBytes, (cudaStream_t)stream>>>(
                                  *p, xh, vec, (const float4 *)geom_g, gptr, (const int4 *)meta, group_rows, group_mod, n_groups,
                                  coef, bias, offset, dx, dvec));
    return hn::check_launch(where);
}

extern "C" int hn_painn_edge_bwd_dst_group(const hn_edge_params *p, const float *xh, const float *vec, const float *geom_g,
                                           const int32_t *gptr, const int32_t *meta, const int32_t *group_rows,
                                           const int32_t *group_mod, int32_t n_groups, const float *coef, const float *bias,
                                           const float *offset, const float *g_dx, const float *g_dvec, float *g_geom_g,
                                           int64_t n_slots, void *stream) {
    const char *where = "hn_painn_edge_bwd_dst_group";
    if (int rc = validate_group(where, p)) return rc;
    if (n_groups <= 0) return 0;
    const long long units = (long long)n_groups * (p->hidden / 64);
    dim3 grid((unsigned)((units + kWarps - 1) / kWarps));
    HN_DISPATCH_NS(p->hidden, HN_CUDA(cudaFuncSetAttribute(edge_bwd_dst_group_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           kWarps * kWarpBytes), where);
                   edge_bwd_dst_group_kernel<NS><<<grid, 32 * kWarps, kWarps * kWarpBytes, (cudaStream_t)stream>>>(
                                  *p, xh, vec, (const float4 *)geom_g, gptr, (const int4 *)meta, group_rows, group_mod, n_groups,
                                  coef, bias, offset, g_dx, g_dvec, g_geom_g, (long long)n_slots));
    return hn::check_launch(where);
}
