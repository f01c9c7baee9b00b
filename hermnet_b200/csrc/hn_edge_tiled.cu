// Tiled ("filter-stationary") PaiNN edge kernels: forward, destination-major backward, source-major backward.
//
// Same arithmetic as hn_edge.cu (rmnet.py:55-73 + rmnet.py:168-193 of the reference), different loop nest.  The
// row-per-warp kernels of hn_edge.cu re-read a 12 x 3F band of the filter matrix for EVERY edge (18 KB of L1
// traffic per edge at F=128) and are bound by the L1/LSU pipe at ~200 clk/edge/SM.  Here the graph builder
// (graph.TilePlan) groups rows into tiles of 16 (same sub-network) and sorts the edges of a tile by the
// 16-wide window [4w, 4w+16) of basis functions that contains their Gaussian band.  One warp owns one
// (tile, 64-channel slice): it keeps the window's filter rows W[4w..4w+15][3][2 channels/lane] in REGISTERS
// (96 registers), streams the bucket's edges through them with packed fp32 FMAs (fma.rn.f32x2 -> FFMA2, two
// channels per instruction) and accumulates the 16 rows' messages in a private shared-memory tile with plain
// read-modify-write -- no atomics, no inter-warp synchronisation, deterministic summation order.
//
// Per edge the basis values env*g_k(d) of the window are evaluated once (one expf per lane, two edges per
// instruction stream: lanes 0-15 / 16-31), duplicated into (g,g) pairs in shared memory and broadcast with
// LDS.128.  Buckets are padded to an even edge count with entries that point at a discarded accumulator row,
// so the pair loop has no tail.  The kernels re-derive every edge's band from the CURRENT distance and fall
// back to a per-edge evaluation when the plan is stale (graph reused after the atoms moved), so results never
// depend on the plan being exact.
#include "hn_common.cuh"

namespace {

typedef unsigned long long u64;

constexpr int kRT = 16;     // rows (or source atoms) per tile
constexpr int kWin = 16;    // basis functions held in registers
constexpr int kStep = 4;    // window stride
constexpr int kBandLo = 5, kBandHi = 6;   // band = floor(x)-5 .. floor(x)+6 (12 terms; dropped terms < e^-18)

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

struct EnvCoef {
    float a, b, c;
    int p;
};
__device__ __forceinline__ EnvCoef env_coef(int p) {
    EnvCoef r;
    r.p = p;
    r.a = -0.5f * (float)((p + 1) * (p + 2));
    r.b = (float)(p * (p + 2));
    r.c = -0.5f * (float)(p * (p + 1));
    return r;
}
__device__ __forceinline__ float ipow(float u, int p) {
    float r = 1.f;
    for (int i = 0; i < p; ++i) r *= u;
    return r;
}
// env(u) and (optionally) d env/du of the polynomial envelope (rmnet.py:183-193)
template <bool DERIV>
__device__ __forceinline__ void envelope(float u, const EnvCoef &E, float &env, float &denv) {
    const float um = E.p == 5 ? (u * u) * (u * u) : ipow(u, E.p - 1), u0 = um * u, u1 = u0 * u, u2 = u1 * u;
    env = 1.f + E.a * u0 + E.b * u1 + E.c * u2;
    if (DERIV) denv = E.a * (float)E.p * um + E.b * (float)(E.p + 1) * u0 + E.c * (float)(E.p + 2) * u1;
}

// window start of bucket w (w < NW) and the test "the 12-term band of distance u lies inside window k0"
__device__ __forceinline__ int window_k0(int w, int K) { return min(kStep * w, K - kWin); }
__device__ __forceinline__ bool band_in_window(float u, int K, int k0) {
    const int kc = (int)floorf(u * (float)(K - 1));
    const int lo = max(kc - kBandLo, 0), hi = min(kc + kBandHi, K - 1);
    return k0 <= lo && hi <= k0 + kWin - 1;
}

// Per-edge fallback (stale plan): phi (and dphi/dd) over the 12-term band straight from global memory.
struct Phi6 {
    u64 fa, fb, fc, da, db, dc;
};
template <bool DERIV>
__device__ __noinline__ Phi6 slow_phi(float u, int K, int env_p, float coeff, float inv_rc, const float *__restrict__ offset,
                                      const float *__restrict__ Wm, int F, u64 ba, u64 bb, u64 bc) {
    Phi6 r;
    r.fa = ba, r.fb = bb, r.fc = bc, r.da = 0ull, r.db = 0ull, r.dc = 0ull;
    if (!(u < 1.f)) return r;
    const int F3 = 3 * F;
    const EnvCoef E = env_coef(env_p);
    float env, denv = 0.f;
    envelope<DERIV>(u, E, env, denv);
    const int kc = (int)floorf(u * (float)(K - 1));
    const int nb = K < 12 ? K : 12;
    int k0 = kc - kBandLo;
    k0 = k0 < 0 ? 0 : k0;
    k0 = k0 > K - nb ? K - nb : k0;
    for (int j = 0; j < nb; ++j) {
        const float diff = u - __ldg(offset + k0 + j);
        const float g = expf(coeff * diff * diff);
        const float val = env * g;
        const float *w = Wm + (size_t)(k0 + j) * F3;
        const u64 wa = ldg64(w), wb = ldg64(w + F), wc = ldg64(w + 2 * F);
        const u64 vv = pk(val, val);
        r.fa = fma2(vv, wa, r.fa);
        r.fb = fma2(vv, wb, r.fb);
        r.fc = fma2(vv, wc, r.fc);
        if (DERIV) {
            const float dval = (denv * g + val * (2.f * coeff * diff)) * inv_rc;
            const u64 dd = pk(dval, dval);
            r.da = fma2(dd, wa, r.da);
            r.db = fma2(dd, wb, r.db);
            r.dc = fma2(dd, wc, r.dc);
        }
    }
    return r;
}

#define HN_LOAD_WINDOW(Wbase, k0)                                                   \
    _Pragma("unroll") for (int jj = 0; jj < kWin; ++jj) {                            \
        const float *wr_ = (Wbase) + (size_t)((k0) + jj) * F3;                       \
        w[jj][0] = ldg64(wr_);                                                       \
        w[jj][1] = ldg64(wr_ + F);                                                   \
        w[jj][2] = ldg64(wr_ + 2 * F);                                               \
    }

// phi over the register window: two independent accumulator sets (even / odd basis function) for ILP
#define HN_CHAIN3(gq, fa, fb, fc)                                                    \
    {                                                                                \
        u64 fa1_ = 0ull, fb1_ = 0ull, fc1_ = 0ull;                                   \
        _Pragma("unroll") for (int jj = 0; jj < kWin / 2; ++jj) {                    \
            const ulonglong2 g2 = (gq)[jj];                                          \
            fa = fma2(g2.x, w[2 * jj][0], fa);                                       \
            fb = fma2(g2.x, w[2 * jj][1], fb);                                       \
            fc = fma2(g2.x, w[2 * jj][2], fc);                                       \
            fa1_ = fma2(g2.y, w[2 * jj + 1][0], fa1_);                               \
            fb1_ = fma2(g2.y, w[2 * jj + 1][1], fb1_);                               \
            fc1_ = fma2(g2.y, w[2 * jj + 1][2], fc1_);                               \
        }                                                                            \
        fa = add2(fa, fa1_);                                                         \
        fb = add2(fb, fb1_);                                                         \
        fc = add2(fc, fc1_);                                                         \
    }

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(unsigned dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kStages = 3;                     // fwd / bwd_dst: pairs in flight (cp.async ring)
constexpr int kStageBytes = 12 * 256 + 64;     // 12 gathered 64-channel slices (2 edges x {xh a,b,c, vec x,y,z}) + header
constexpr int kAccDst = (kRT + 1) * 4 * 64;    // floats per warp: rows x {dx, dvec_x, dvec_y, dvec_z} x 64 channels
constexpr int kWarpsDst = 4;

// The slot header of a pair: lanes 0..3 fetch {meta[e], meta[e+1], geom[e], geom[e+1]} (16 B each).
__device__ __forceinline__ int4 load_piece(const int4 *meta, const float4 *geom, int e, bool valid, int lane) {
    int4 r = make_int4(0, 0, 0, 0);
    if (valid && lane < 4) r = __ldg(lane < 2 ? meta + e + lane : reinterpret_cast<const int4 *>(geom + e + (lane - 2)));
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// Software pipeline per warp, pair p = slots (E0+2p, E0+2p+1):
//   iteration p:  cp.async gathers of pair p+2 | basis values of pair p+1 (expf) | filter chains + messages of pair p
// ---------------------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(32 * kWarpsDst, 2)
edge_fwd_tiled_kernel(const hn_edge_params P, const float *__restrict__ xh, const float *__restrict__ vec,
                      const float4 *__restrict__ geom_b, const int *__restrict__ bptr, const int4 *__restrict__ meta,
                      const int *__restrict__ tile_rows, const int *__restrict__ tile_mod, int n_tiles, int NW,
                      const float *__restrict__ Wt, const float *__restrict__ bias, const float *__restrict__ offset,
                      float *__restrict__ dx, float *__restrict__ dvec) {
    extern __shared__ __align__(16) float smem[];
    constexpr int F = 64 * NS, F3 = 3 * F;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kWarpsDst + warp;
    const int tile = unit / NS, slice = unit - tile * NS;
    if (tile >= n_tiles) return;
    float *acc = smem + warp * kAccDst;
    u64 *gbuf = reinterpret_cast<u64 *>(smem + kWarpsDst * kAccDst) + warp * 64;                  // [2 pairs][2 edges][16] (g,g)
    char *stg = reinterpret_cast<char *>(smem + kWarpsDst * kAccDst) + kWarpsDst * 512 + warp * (kStages * kStageBytes);
    for (int i = lane; i < kAccDst / 2; i += 32) reinterpret_cast<u64 *>(acc)[i] = 0ull;
    const int m = __ldg(tile_mod + tile);
    const int K = P.num_rbf;
    const int ch = slice * 64 + 2 * lane;
    const float *Wm = Wt + (size_t)m * K * F3 + ch;
    const u64 ba = ldg64(bias + (size_t)m * F3 + ch), bb = ldg64(bias + (size_t)m * F3 + F + ch),
              bc = ldg64(bias + (size_t)m * F3 + 2 * F + ch);
    const float c1 = 1.0f / sqrtf(3.0f * (float)F), c2 = 1.0f / sqrtf((float)F);
    const u64 c1p = pk(c1, c1), c2p = pk(c2, c2);
    const EnvCoef E = env_coef(P.env_p);
    const int *bp = bptr + (size_t)tile * (NW + 1);
    const bool bp_in_lanes = NW + 2 <= 32;
    const int bpl = (bp_in_lanes && lane <= NW + 1) ? __ldg(bp + lane) : 0;
    const int E0 = __ldg(bp), E1 = __ldg(bp + NW + 1);
    const int np = (E1 - E0) >> 1;
    const int h = lane >> 4, j = lane & 15;
    u64 w[kWin][3];
#pragma unroll
    for (int jj = 0; jj < kWin; ++jj) w[jj][0] = w[jj][1] = w[jj][2] = 0ull;
    __syncwarp();
    if (np == 0) goto write_out;
    {
        auto issue = [&](char *hdr, bool valid, const int4 &piece) {
            if (lane < 4) *reinterpret_cast<int4 *>(hdr + lane * 16) = piece;
            const int srcA = __shfl_sync(0xffffffffu, piece.x, 0), xrA = __shfl_sync(0xffffffffu, piece.z, 0);
            const int srcB = __shfl_sync(0xffffffffu, piece.x, 1), xrB = __shfl_sync(0xffffffffu, piece.z, 1);
            if (valid) {
                const float *xa = xh + (size_t)xrA * F3 + ch, *va = vec + (size_t)srcA * F3 + ch;
                const float *xb = xh + (size_t)xrB * F3 + ch, *vb = vec + (size_t)srcB * F3 + ch;
                const unsigned d = smem_u32(hdr + 64) + lane * 8;
                cp_async8(d, xa);
                cp_async8(d + 256, xa + F);
                cp_async8(d + 512, xa + 2 * F);
                cp_async8(d + 768, va);
                cp_async8(d + 1024, va + F);
                cp_async8(d + 1280, va + 2 * F);
                cp_async8(d + 1536, xb);
                cp_async8(d + 1792, xb + F);
                cp_async8(d + 2048, xb + 2 * F);
                cp_async8(d + 2304, vb);
                cp_async8(d + 2560, vb + F);
                cp_async8(d + 2816, vb + 2 * F);
            }
            cp_commit();
        };
        // window state of the pair whose basis values are being evaluated
        int wi = -1, wend = E0, k0 = 0;
        bool in_range = false;
        float off = 0.f;
        // basis values of the pair starting at slot e (header hdr) -> gbuf half `buf`; returns (ok mask, window changed)
        auto basis = [&](int e, const char *hdr, int buf, bool &changed) -> unsigned {
            changed = false;
            if (e >= wend) {
                do {
                    ++wi;
                    wend = bp_in_lanes ? __shfl_sync(0xffffffffu, bpl, wi + 1) : __ldg(bp + wi + 1);
                } while (e >= wend);
                in_range = wi < NW;
                k0 = in_range ? window_k0(wi, K) : 0;
                off = __ldg(offset + k0 + j);
                changed = true;
            }
            const float d = *reinterpret_cast<const float *>(hdr + 32 + 16 * h + 12);
            const float u = d * P.inv_rc;
            float val = 0.f;
            bool ok = true;
            if (u < 1.f) {
                ok = in_range && band_in_window(u, K, k0);
                float env, denv;
                envelope<false>(u, E, env, denv);
                const float diff = u - off;
                val = in_range ? env * expf(P.coeff * diff * diff) : 0.f;
            }
            gbuf[buf * 32 + lane] = pk(val, val);
            return __ballot_sync(0xffffffffu, ok);
        };

        char *st0 = stg, *st1 = stg + kStageBytes, *st2 = stg + 2 * kStageBytes;   // stages of pairs p, p+1, p+2
        issue(st0, true, load_piece(meta, geom_b, E0, true, lane));
        issue(st1, 1 < np, load_piece(meta, geom_b, E0 + 2, 1 < np, lane));
        int4 pnext = load_piece(meta, geom_b, E0 + 4, 2 < np, lane);
        __syncwarp();
        bool changed;
        unsigned okm = basis(E0, st0, 0, changed);
        int k0c = k0;          // window of the pair being consumed
        if (in_range) { HN_LOAD_WINDOW(Wm, k0c) }
        for (int p = 0; p < np; ++p) {
            const int4 pnn = load_piece(meta, geom_b, E0 + 2 * (p + 3), p + 3 < np, lane);
            issue(st2, p + 2 < np, pnext);
            pnext = pnn;
            cp_wait<2>();
            __syncwarp();                       // header of pair p+1, basis values of pair p visible
            unsigned okm_n = 0xffffffffu;
            bool changed_n = false;
            if (p + 1 < np) okm_n = basis(E0 + 2 * (p + 1), st1, (p + 1) & 1, changed_n);
            const int4 mA = *reinterpret_cast<const int4 *>(st0), mB = *reinterpret_cast<const int4 *>(st0 + 16);
            const float4 gA = *reinterpret_cast<const float4 *>(st0 + 32), gB = *reinterpret_cast<const float4 *>(st0 + 48);
            const u64 *dq = reinterpret_cast<const u64 *>(st0 + 64) + lane;
            u64 faA = ba, fbA = bb, fcA = bc, faB = ba, fbB = bb, fcB = bc;
            if (okm == 0xffffffffu) {
                const ulonglong2 *gq = reinterpret_cast<const ulonglong2 *>(gbuf + (p & 1) * 32);
#pragma unroll
                for (int jj = 0; jj < kWin / 2; ++jj) {
                    const ulonglong2 a2 = gq[jj], b2 = gq[8 + jj];
                    faA = fma2(a2.x, w[2 * jj][0], faA);
                    fbA = fma2(a2.x, w[2 * jj][1], fbA);
                    fcA = fma2(a2.x, w[2 * jj][2], fcA);
                    faB = fma2(b2.x, w[2 * jj][0], faB);
                    fbB = fma2(b2.x, w[2 * jj][1], fbB);
                    fcB = fma2(b2.x, w[2 * jj][2], fcB);
                    faA = fma2(a2.y, w[2 * jj + 1][0], faA);
                    fbA = fma2(a2.y, w[2 * jj + 1][1], fbA);
                    fcA = fma2(a2.y, w[2 * jj + 1][2], fcA);
                    faB = fma2(b2.y, w[2 * jj + 1][0], faB);
                    fbB = fma2(b2.y, w[2 * jj + 1][1], fbB);
                    fcB = fma2(b2.y, w[2 * jj + 1][2], fcB);
                }
            } else {                            // stale plan: at least one band left its window
                const Phi6 a = slow_phi<false>(gA.w * P.inv_rc, K, P.env_p, P.coeff, P.inv_rc, offset, Wm, F, ba, bb, bc);
                const Phi6 b = slow_phi<false>(gB.w * P.inv_rc, K, P.env_p, P.coeff, P.inv_rc, offset, Wm, F, ba, bb, bc);
                faA = a.fa, fbA = a.fb, fcA = a.fc, faB = b.fa, fbB = b.fb, fcB = b.fc;
            }
            {
                const u64 Pa = dq[0], Pb = dq[32], Pc = dq[64], V0 = dq[96], V1 = dq[128], V2 = dq[160];
                u64 *ar = reinterpret_cast<u64 *>(acc + mA.y * 256) + lane;
                const u64 tb = mul2(mul2(Pb, fbA), c1p);
                const u64 tc = mul2(mul2(Pc, fcA), c2p);
                ar[0] = fma2(Pa, faA, ar[0]);
                ar[32] = fma2(tc, pk(gA.x, gA.x), fma2(V0, tb, ar[32]));
                ar[64] = fma2(tc, pk(gA.y, gA.y), fma2(V1, tb, ar[64]));
                ar[96] = fma2(tc, pk(gA.z, gA.z), fma2(V2, tb, ar[96]));
            }
            {
                const u64 Pa = dq[192], Pb = dq[224], Pc = dq[256], V0 = dq[288], V1 = dq[320], V2 = dq[352];
                u64 *ar = reinterpret_cast<u64 *>(acc + mB.y * 256) + lane;
                const u64 tb = mul2(mul2(Pb, fbB), c1p);
                const u64 tc = mul2(mul2(Pc, fcB), c2p);
                ar[0] = fma2(Pa, faB, ar[0]);
                ar[32] = fma2(tc, pk(gB.x, gB.x), fma2(V0, tb, ar[32]));
                ar[64] = fma2(tc, pk(gB.y, gB.y), fma2(V1, tb, ar[64]));
                ar[96] = fma2(tc, pk(gB.z, gB.z), fma2(V2, tb, ar[96]));
            }
            if (changed_n && in_range) { HN_LOAD_WINDOW(Wm, k0) }   // filter rows of the next pair's window
            okm = okm_n;
            char *t = st0;
            st0 = st1;
            st1 = st2;
            st2 = t;
            __syncwarp();
        }
        cp_wait<0>();
        __syncwarp();
    }
write_out:
    for (int r = 0; r < kRT; ++r) {
        const int row = __ldg(tile_rows + (size_t)tile * kRT + r);
        if (row < 0) continue;
        const u64 *ar = reinterpret_cast<const u64 *>(acc + r * 256) + lane;
        *reinterpret_cast<u64 *>(dx + (size_t)row * F + ch) = ar[0];
        *reinterpret_cast<u64 *>(dvec + (size_t)row * F3 + ch) = ar[32];
        *reinterpret_cast<u64 *>(dvec + (size_t)row * F3 + F + ch) = ar[64];
        *reinterpret_cast<u64 *>(dvec + (size_t)row * F3 + 2 * F + ch) = ar[96];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// destination-major backward: per edge (dL/du_x, dL/du_y, dL/du_z, dL/dd), partial over this 64-channel slice,
// written in SLOT order (the caller adds the slices and un-permutes).
// ---------------------------------------------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(32 * kWarpsDst, 2)
edge_bwd_dst_tiled_kernel(const hn_edge_params P, const float *__restrict__ xh, const float *__restrict__ vec,
                          const float4 *__restrict__ geom_b, const int *__restrict__ bptr, const int4 *__restrict__ meta,
                          const int *__restrict__ tile_rows, const int *__restrict__ tile_mod, int n_tiles, int NW,
                          const float *__restrict__ Wt, const float *__restrict__ bias, const float *__restrict__ offset,
                          const float *__restrict__ g_dx, const float *__restrict__ g_dvec, float *__restrict__ g_geom_b,
                          long long n_pad) {
    extern __shared__ __align__(16) float smem[];
    constexpr int F = 64 * NS, F3 = 3 * F;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kWarpsDst + warp;
    const int tile = unit / NS, slice = unit - tile * NS;
    if (tile >= n_tiles) return;
    float *gs = smem + warp * kAccDst;                                                            // upstream gradients
    u64 *gbuf = reinterpret_cast<u64 *>(smem + kWarpsDst * kAccDst) + warp * 64;                  // (g,g) x32 then (g',g') x32
    char *stg = reinterpret_cast<char *>(smem + kWarpsDst * kAccDst) + kWarpsDst * 512 + warp * (kStages * kStageBytes);
    const int ch = slice * 64 + 2 * lane;
    for (int r = 0; r <= kRT; ++r) {
        const int row = r < kRT ? __ldg(tile_rows + (size_t)tile * kRT + r) : -1;
        u64 *gr = reinterpret_cast<u64 *>(gs + r * 256) + lane;
        if (row >= 0) {
            gr[0] = ldg64(g_dx + (size_t)row * F + ch);
            gr[32] = ldg64(g_dvec + (size_t)row * F3 + ch);
            gr[64] = ldg64(g_dvec + (size_t)row * F3 + F + ch);
            gr[96] = ldg64(g_dvec + (size_t)row * F3 + 2 * F + ch);
        } else {
            gr[0] = gr[32] = gr[64] = gr[96] = 0ull;
        }
    }
    const int m = __ldg(tile_mod + tile);
    const int K = P.num_rbf;
    const float *Wm = Wt + (size_t)m * K * F3 + ch;
    const u64 bc = ldg64(bias + (size_t)m * F3 + 2 * F + ch);
    const float c1 = 1.0f / sqrtf(3.0f * (float)F), c2 = 1.0f / sqrtf((float)F);
    const u64 c1p = pk(c1, c1), c2p = pk(c2, c2);
    const EnvCoef E = env_coef(P.env_p);
    const int *bp = bptr + (size_t)tile * (NW + 1);
    const bool bp_in_lanes = NW + 2 <= 32;
    const int bpl = (bp_in_lanes && lane <= NW + 1) ? __ldg(bp + lane) : 0;
    const int E0 = __ldg(bp), E1 = __ldg(bp + NW + 1);
    const int np = (E1 - E0) >> 1;
    const int h = lane >> 4, j = lane & 15;
    float *out = g_geom_b + (size_t)slice * n_pad * 4;
    u64 w[kWin][3];
#pragma unroll
    for (int jj = 0; jj < kWin; ++jj) w[jj][0] = w[jj][1] = w[jj][2] = 0ull;
    __syncwarp();

    auto issue = [&](int p, const int4 &piece) {
        int st = p % kStages;
        char *hdr = stg + st * kStageBytes;
        if (lane < 4) *reinterpret_cast<int4 *>(hdr + lane * 16) = piece;
        const int srcA = __shfl_sync(0xffffffffu, piece.x, 0), xrA = __shfl_sync(0xffffffffu, piece.z, 0);
        const int srcB = __shfl_sync(0xffffffffu, piece.x, 1), xrB = __shfl_sync(0xffffffffu, piece.z, 1);
        if (p < np) {
            const float *xa = xh + (size_t)xrA * F3 + ch, *va = vec + (size_t)srcA * F3 + ch;
            const float *xb = xh + (size_t)xrB * F3 + ch, *vb = vec + (size_t)srcB * F3 + ch;
            const unsigned d = smem_u32(hdr + 64) + lane * 8;
            cp_async8(d, xa);
            cp_async8(d + 256, xa + F);
            cp_async8(d + 512, xa + 2 * F);
            cp_async8(d + 768, va);
            cp_async8(d + 1024, va + F);
            cp_async8(d + 1280, va + 2 * F);
            cp_async8(d + 1536, xb);
            cp_async8(d + 1792, xb + F);
            cp_async8(d + 2048, xb + 2 * F);
            cp_async8(d + 2304, vb);
            cp_async8(d + 2560, vb + F);
            cp_async8(d + 2816, vb + 2 * F);
        }
        cp_commit();
    };

    {
        const int4 p0 = load_piece(meta, geom_b, E0, 0 < np, lane), p1 = load_piece(meta, geom_b, E0 + 2, 1 < np, lane);
        issue(0, p0);
        issue(1, p1);
    }
    int4 pnext = load_piece(meta, geom_b, E0 + 4, 2 < np, lane);
    int wi = -1, wend = E0, k0 = 0;
    bool in_range = false;
    float off = 0.f;
    for (int p = 0; p < np; ++p) {
        const int4 pnn = load_piece(meta, geom_b, E0 + 2 * (p + 3), p + 3 < np, lane);
        issue(p + 2, pnext);
        pnext = pnn;
        cp_wait<2>();
        __syncwarp();
        const char *hdr = stg + (p % kStages) * kStageBytes;
        const int4 mA = *reinterpret_cast<const int4 *>(hdr), mB = *reinterpret_cast<const int4 *>(hdr + 16);
        const float4 gA = *reinterpret_cast<const float4 *>(hdr + 32), gB = *reinterpret_cast<const float4 *>(hdr + 48);
        const u64 *dq = reinterpret_cast<const u64 *>(hdr + 64) + lane;
        const int e = E0 + 2 * p;
        if (e >= wend) {
            do {
                ++wi;
                wend = bp_in_lanes ? __shfl_sync(0xffffffffu, bpl, wi + 1) : __ldg(bp + wi + 1);
            } while (e >= wend);
            in_range = wi < NW;
            k0 = 0;
            if (in_range) {
                k0 = window_k0(wi, K);
                HN_LOAD_WINDOW(Wm, k0)
            }
            off = __ldg(offset + k0 + j);
        }
        unsigned okm;
        {
            const float u = (h ? gB.w : gA.w) * P.inv_rc;
            float val = 0.f, dval = 0.f;
            bool ok = true;
            if (u < 1.f) {
                ok = in_range && band_in_window(u, K, k0);
                if (in_range) {
                    float env, denv;
                    envelope<true>(u, E, env, denv);
                    const float diff = u - off;
                    const float g = expf(P.coeff * diff * diff);
                    val = env * g;
                    dval = (denv * g + val * (2.f * P.coeff * diff)) * P.inv_rc;
                }
            }
            gbuf[lane] = pk(val, val);
            gbuf[32 + lane] = pk(dval, dval);
            okm = __ballot_sync(0xffffffffu, ok);
        }
        __syncwarp();
        float r8[8];
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
            const int4 mX = sub ? mB : mA;
            const float4 gX = sub ? gB : gA;
            const u64 Pa = dq[(sub * 6 + 0) * 32], Pb = dq[(sub * 6 + 1) * 32], Pc = dq[(sub * 6 + 2) * 32];
            const u64 V0 = dq[(sub * 6 + 3) * 32], V1 = dq[(sub * 6 + 4) * 32], V2 = dq[(sub * 6 + 5) * 32];
            u64 fa = 0ull, fb = 0ull, fc = bc, da = 0ull, db = 0ull, dc = 0ull;
            if (!((okm >> (16 * sub)) & 1u)) {
                const Phi6 r = slow_phi<true>(gX.w * P.inv_rc, K, P.env_p, P.coeff, P.inv_rc, offset, Wm, F, 0ull, 0ull, bc);
                fa = r.fa, fb = r.fb, fc = r.fc, da = r.da, db = r.db, dc = r.dc;
            } else if (in_range) {
                const ulonglong2 *gq = reinterpret_cast<const ulonglong2 *>(gbuf + sub * 16);
                const ulonglong2 *hq = reinterpret_cast<const ulonglong2 *>(gbuf + 32 + sub * 16);
                u64 fc1 = 0ull, da1 = 0ull, db1 = 0ull, dc1 = 0ull;
#pragma unroll
                for (int jj = 0; jj < kWin / 2; ++jj) {
                    const ulonglong2 g2 = gq[jj], h2 = hq[jj];
                    fc = fma2(g2.x, w[2 * jj][2], fc);
                    da = fma2(h2.x, w[2 * jj][0], da);
                    db = fma2(h2.x, w[2 * jj][1], db);
                    dc = fma2(h2.x, w[2 * jj][2], dc);
                    fc1 = fma2(g2.y, w[2 * jj + 1][2], fc1);
                    da1 = fma2(h2.y, w[2 * jj + 1][0], da1);
                    db1 = fma2(h2.y, w[2 * jj + 1][1], db1);
                    dc1 = fma2(h2.y, w[2 * jj + 1][2], dc1);
                }
                fc = add2(fc, fc1);
                da = add2(da, da1);
                db = add2(db, db1);
                dc = add2(dc, dc1);
            }
            const u64 *gr = reinterpret_cast<const u64 *>(gs + mX.y * 256) + lane;
            const u64 gx = gr[0], gv0 = gr[32], gv1 = gr[64], gv2 = gr[96];
            const u64 tb = mul2(fma2(gv2, V2, fma2(gv1, V1, mul2(gv0, V0))), c1p);                          // dL/d(Pb*phib)
            const u64 tc = mul2(fma2(gv2, pk(gX.z, gX.z), fma2(gv1, pk(gX.y, gX.y), mul2(gv0, pk(gX.x, gX.x)))), c2p);
            const u64 gd = fma2(mul2(tc, Pc), dc, fma2(mul2(tb, Pb), db, mul2(mul2(gx, Pa), da)));
            const u64 cphi = mul2(mul2(Pc, fc), c2p);
            r8[4 * sub + 0] = hsum(mul2(gv0, cphi));
            r8[4 * sub + 1] = hsum(mul2(gv1, cphi));
            r8[4 * sub + 2] = hsum(mul2(gv2, cphi));
            r8[4 * sub + 3] = hsum(gd);
        }
        // 8 values x 32 lanes -> 8 sums with 9 shuffles (halving butterfly)
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        float a4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = b4 ? r8[i] : r8[4 + i];
            const float keep = b4 ? r8[4 + i] : r8[i];
            a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        float a2[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = b3 ? a4[i] : a4[2 + i];
            const float keep = b3 ? a4[2 + i] : a4[i];
            a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        float a1;
        {
            const float send = b2 ? a2[0] : a2[1];
            const float keep = b2 ? a2[1] : a2[0];
            a1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        if ((lane & 3) == 0) out[(size_t)(e + (b4 ? 1 : 0)) * 4 + (b3 ? 2 : 0) + (b2 ? 1 : 0)] = a1;
        __syncwarp();
    }
    cp_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------------
// source-major backward over the transposed plan: tiles of 16 SOURCE atoms, buckets keyed (module, window).
// grad_vec[s] is accumulated over all modules; grad_xh[(m,s)] per module, flushed when the module changes
// (each (m,s) row of the zero-filled grad_xh buffer is written by exactly one warp).
// meta = (destination row, local source, xh row of the source under this module, -)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWarpsSrc = 3;
constexpr int kStagesSrc = 2;
constexpr int kStageBytesSrc = 16 * 256 + 64;   // 2 edges x {g_dx, g_dvec x,y,z, xh_b, vec x,y,z}
constexpr int kAccSrc = (kRT + 1) * 3 * 64;     // floats: sources x 3 parts x 64 channels
constexpr int kWarpBytesSrc = 2 * kAccSrc * 4 + 160 + 256 + kStagesSrc * kStageBytesSrc;   // acc | xrow | gbuf | stages

template <int NS>
__global__ void __launch_bounds__(32 * kWarpsSrc, 2)
edge_bwd_src_tiled_kernel(const hn_edge_params P, const float *__restrict__ xh, const float *__restrict__ vec,
                          const float4 *__restrict__ geom_s, const int *__restrict__ bptr, const int4 *__restrict__ meta,
                          int n_tiles, int NW, const float *__restrict__ Wt, const float *__restrict__ bias,
                          const float *__restrict__ offset, const float *__restrict__ g_dx,
                          const float *__restrict__ g_dvec, float *__restrict__ grad_xh, float *__restrict__ grad_vec) {
    extern __shared__ __align__(16) float smem[];
    constexpr int F = 64 * NS, F3 = 3 * F;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kWarpsSrc + warp;
    const int tile = unit / NS, slice = unit - tile * NS;
    if (tile >= n_tiles) return;
    char *base = reinterpret_cast<char *>(smem) + (size_t)warp * kWarpBytesSrc;
    float *accv = reinterpret_cast<float *>(base);
    float *accx = accv + kAccSrc;
    int *xrow = reinterpret_cast<int *>(base + 2 * kAccSrc * 4);
    u64 *gbuf = reinterpret_cast<u64 *>(base + 2 * kAccSrc * 4 + 160);
    char *stg = base + 2 * kAccSrc * 4 + 160 + 256;
    for (int i = lane; i < kAccSrc; i += 32) reinterpret_cast<u64 *>(accv)[i] = 0ull;   // accv and accx
    if (lane <= kRT) xrow[lane] = -1;
    const int K = P.num_rbf, M = P.n_modules, NB = NW + 1;
    const int ch = slice * 64 + 2 * lane;
    const float c1 = 1.0f / sqrtf(3.0f * (float)F), c2 = 1.0f / sqrtf((float)F);
    const u64 c1p = pk(c1, c1), c2p = pk(c2, c2);
    const EnvCoef E = env_coef(P.env_p);
    const int h = lane >> 4, j = lane & 15;
    const int s0 = tile * kRT;
    const int *bp = bptr + (size_t)tile * M * NB;
    const int E0 = __ldg(bp), E1 = __ldg(bp + M * NB);
    const int np = (E1 - E0) >> 1;
    u64 w[kWin][3];
#pragma unroll
    for (int jj = 0; jj < kWin; ++jj) w[jj][0] = w[jj][1] = w[jj][2] = 0ull;
    __syncwarp();

    auto issue = [&](int p, const int4 &piece) {
        int st = p % kStagesSrc;
        char *hdr = stg + st * kStageBytesSrc;
        if (lane < 4) *reinterpret_cast<int4 *>(hdr + lane * 16) = piece;
        const int rowA = __shfl_sync(0xffffffffu, piece.x, 0), lsA = __shfl_sync(0xffffffffu, piece.y, 0),
                  xrA = __shfl_sync(0xffffffffu, piece.z, 0);
        const int rowB = __shfl_sync(0xffffffffu, piece.x, 1), lsB = __shfl_sync(0xffffffffu, piece.y, 1),
                  xrB = __shfl_sync(0xffffffffu, piece.z, 1);
        if (p < np) {
            const int sA = min(s0 + min(lsA, kRT - 1), P.n_atoms - 1), sB = min(s0 + min(lsB, kRT - 1), P.n_atoms - 1);
            const float *ga = g_dvec + (size_t)rowA * F3 + ch, *gb = g_dvec + (size_t)rowB * F3 + ch;
            const float *va = vec + (size_t)sA * F3 + ch, *vb = vec + (size_t)sB * F3 + ch;
            const unsigned d = smem_u32(hdr + 64) + lane * 8;
            cp_async8(d, g_dx + (size_t)rowA * F + ch);
            cp_async8(d + 256, ga);
            cp_async8(d + 512, ga + F);
            cp_async8(d + 768, ga + 2 * F);
            cp_async8(d + 1024, xh + (size_t)xrA * F3 + F + ch);
            cp_async8(d + 1280, va);
            cp_async8(d + 1536, va + F);
            cp_async8(d + 1792, va + 2 * F);
            cp_async8(d + 2048, g_dx + (size_t)rowB * F + ch);
            cp_async8(d + 2304, gb);
            cp_async8(d + 2560, gb + F);
            cp_async8(d + 2816, gb + 2 * F);
            cp_async8(d + 3072, xh + (size_t)xrB * F3 + F + ch);
            cp_async8(d + 3328, vb);
            cp_async8(d + 3584, vb + F);
            cp_async8(d + 3840, vb + 2 * F);
        }
        cp_commit();
    };

    auto flush = [&]() {      // write this module's grad_xh rows, clear the accumulator and the row table
        __syncwarp();
        for (int r = 0; r <= kRT; ++r) {
            const int xr = r < kRT ? xrow[r] : -1;
            u64 *ax = reinterpret_cast<u64 *>(accx + r * 192) + lane;
            if (xr >= 0) {
                float *dst = grad_xh + (size_t)xr * F3 + ch;
                *reinterpret_cast<u64 *>(dst) = ax[0];
                *reinterpret_cast<u64 *>(dst + F) = ax[32];
                *reinterpret_cast<u64 *>(dst + 2 * F) = ax[64];
            }
            ax[0] = ax[32] = ax[64] = 0ull;
        }
        __syncwarp();
        if (lane <= kRT) xrow[lane] = -1;
        __syncwarp();
    };

    issue(0, load_piece(meta, geom_s, E0, 0 < np, lane));
    int4 pnext = load_piece(meta, geom_s, E0 + 2, 1 < np, lane);
    int b = -1, bend = E0, k0 = 0, m = -1;
    bool in_range = false;
    float off = 0.f;
    const float *Wm = Wt + ch;
    u64 ba = 0ull, bb = 0ull, bc = 0ull;
    for (int p = 0; p < np; ++p) {
        const int4 pnn = load_piece(meta, geom_s, E0 + 2 * (p + 2), p + 2 < np, lane);
        issue(p + 1, pnext);
        pnext = pnn;
        cp_wait<1>();
        __syncwarp();
        const char *hdr = stg + (p % kStagesSrc) * kStageBytesSrc;
        const int4 mA = *reinterpret_cast<const int4 *>(hdr), mB = *reinterpret_cast<const int4 *>(hdr + 16);
        const float4 gA = *reinterpret_cast<const float4 *>(hdr + 32), gB = *reinterpret_cast<const float4 *>(hdr + 48);
        const u64 *dq = reinterpret_cast<const u64 *>(hdr + 64) + lane;
        const int e = E0 + 2 * p;
        if (e >= bend) {
            do {
                ++b;
                bend = __ldg(bp + b + 1);
            } while (e >= bend);
            const int m_new = b / NB, wi = b - m_new * NB;
            if (m_new != m) {
                if (m >= 0) flush();
                m = m_new;
                Wm = Wt + (size_t)m * K * F3 + ch;
                ba = ldg64(bias + (size_t)m * F3 + ch);
                bb = ldg64(bias + (size_t)m * F3 + F + ch);
                bc = ldg64(bias + (size_t)m * F3 + 2 * F + ch);
            }
            in_range = wi < NW;
            k0 = 0;
            if (in_range) {
                k0 = window_k0(wi, K);
                HN_LOAD_WINDOW(Wm, k0)
            }
            off = __ldg(offset + k0 + j);
        }
        unsigned okm;
        {
            const float u = (h ? gB.w : gA.w) * P.inv_rc;
            float val = 0.f;
            bool ok = true;
            if (u < 1.f) {
                ok = in_range && band_in_window(u, K, k0);
                if (in_range) {
                    float env, denv;
                    envelope<false>(u, E, env, denv);
                    const float diff = u - off;
                    val = env * expf(P.coeff * diff * diff);
                }
            }
            gbuf[lane] = pk(val, val);
            okm = __ballot_sync(0xffffffffu, ok);
        }
        if (lane == 0) {
            xrow[mA.y] = mA.z;
            xrow[mB.y] = mB.z;
        }
        __syncwarp();
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
            const int4 mX = sub ? mB : mA;
            const float4 gX = sub ? gB : gA;
            const u64 gx = dq[(sub * 8 + 0) * 32], gv0 = dq[(sub * 8 + 1) * 32], gv1 = dq[(sub * 8 + 2) * 32],
                      gv2 = dq[(sub * 8 + 3) * 32];
            const u64 Pb = dq[(sub * 8 + 4) * 32], V0 = dq[(sub * 8 + 5) * 32], V1 = dq[(sub * 8 + 6) * 32],
                      V2 = dq[(sub * 8 + 7) * 32];
            u64 fa = ba, fb = bb, fc = bc;
            if (!((okm >> (16 * sub)) & 1u)) {
                const Phi6 r = slow_phi<false>(gX.w * P.inv_rc, K, P.env_p, P.coeff, P.inv_rc, offset, Wm, F, ba, bb, bc);
                fa = r.fa, fb = r.fb, fc = r.fc;
            } else if (in_range) {
                const ulonglong2 *gq = reinterpret_cast<const ulonglong2 *>(gbuf + sub * 16);
                HN_CHAIN3(gq, fa, fb, fc)
            }
            const u64 tb = mul2(fma2(gv2, V2, fma2(gv1, V1, mul2(gv0, V0))), c1p);
            const u64 tc = mul2(fma2(gv2, pk(gX.z, gX.z), fma2(gv1, pk(gX.y, gX.y), mul2(gv0, pk(gX.x, gX.x)))), c2p);
            u64 *ax = reinterpret_cast<u64 *>(accx + mX.y * 192) + lane;
            u64 *av = reinterpret_cast<u64 *>(accv + mX.y * 192) + lane;
            ax[0] = fma2(gx, fa, ax[0]);
            ax[32] = fma2(tb, fb, ax[32]);
            ax[64] = fma2(tc, fc, ax[64]);
            const u64 bphi = mul2(mul2(Pb, fb), c1p);
            av[0] = fma2(gv0, bphi, av[0]);
            av[32] = fma2(gv1, bphi, av[32]);
            av[64] = fma2(gv2, bphi, av[64]);
        }
        __syncwarp();
    }
    cp_wait<0>();
    if (m >= 0) flush();
    for (int r = 0; r < kRT; ++r) {
        const int s = s0 + r;
        if (s >= P.n_atoms) break;
        const u64 *av = reinterpret_cast<const u64 *>(accv + r * 192) + lane;
        float *dst = grad_vec + (size_t)s * F3 + ch;
        *reinterpret_cast<u64 *>(dst) = av[0];
        *reinterpret_cast<u64 *>(dst + F) = av[32];
        *reinterpret_cast<u64 *>(dst + 2 * F) = av[64];
    }
}

int validate_tiled(const char *where, const hn_edge_params *p, int n_windows) {
    HN_REQUIRE(p != nullptr, where, "null params");
    HN_REQUIRE(p->hidden % 64 == 0 && p->hidden >= 64 && p->hidden <= 512, where,
               "tiled edge kernels need hidden_channels in {64,128,...,512}");
    HN_REQUIRE(p->num_rbf >= kWin, where, "tiled edge kernels need num_rbf >= 16");
    HN_REQUIRE(p->env_p >= 1, where, "envelope exponent must be >= 1");
    HN_REQUIRE(p->n_modules >= 1, where, "n_modules must be >= 1");
    HN_REQUIRE(n_windows == (p->num_rbf - kWin + kStep - 1) / kStep + 1, where, "n_windows does not match num_rbf");
    return 0;
}

template <typename Kern>
int set_smem(Kern kern, size_t bytes, const char *where) {
    HN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), where);
    return 0;
}

}  // namespace

#define HN_DISPATCH_NS(F, ...)                                    \
    switch ((F) / 64) {                                           \
        case 1: { constexpr int NS = 1; __VA_ARGS__; break; }            \
        case 2: { constexpr int NS = 2; __VA_ARGS__; break; }            \
        case 3: { constexpr int NS = 3; __VA_ARGS__; break; }            \
        case 4: { constexpr int NS = 4; __VA_ARGS__; break; }            \
        case 5: { constexpr int NS = 5; __VA_ARGS__; break; }            \
        case 6: { constexpr int NS = 6; __VA_ARGS__; break; }            \
        case 7: { constexpr int NS = 7; __VA_ARGS__; break; }            \
        default: { constexpr int NS = 8; __VA_ARGS__; break; }           \
    }

extern "C" int32_t hn_painn_edge_tiled_supported(int32_t hidden, int32_t num_rbf) {
    return (hidden % 64 == 0 && hidden >= 64 && hidden <= 512 && num_rbf >= kWin) ? 1 : 0;
}

extern "C" int32_t hn_painn_edge_tiled_windows(int32_t num_rbf) {
    return num_rbf < kWin ? 0 : (num_rbf - kWin + kStep - 1) / kStep + 1;
}

extern "C" int hn_painn_edge_fwd_tiled(const hn_edge_params *p, const float *xh, const float *vec, const float *geom_b,
                                       const int32_t *bptr, const int32_t *meta, const int32_t *tile_rows,
                                       const int32_t *tile_mod, int32_t n_tiles, int32_t n_windows, const float *Wt,
                                       const float *bias, const float *offset, float *dx, float *dvec, void *stream) {
    const char *where = "hn_painn_edge_fwd_tiled";
    if (int rc = validate_tiled(where, p, n_windows)) return rc;
    if (n_tiles <= 0) return 0;
    const size_t smem = kWarpsDst * (kAccDst * sizeof(float) + 512 + kStages * kStageBytes);
    const int ns = p->hidden / 64;
    const long long units = (long long)n_tiles * ns;
    dim3 grid((unsigned)((units + kWarpsDst - 1) / kWarpsDst));
    HN_DISPATCH_NS(p->hidden, {
        if (int rc = set_smem(edge_fwd_tiled_kernel<NS>, smem, where)) return rc;
        edge_fwd_tiled_kernel<NS><<<grid, 32 * kWarpsDst, smem, (cudaStream_t)stream>>>(
            *p, xh, vec, (const float4 *)geom_b, bptr, (const int4 *)meta, tile_rows, tile_mod, n_tiles, n_windows, Wt, bias,
            offset, dx, dvec);
    });
    return hn::check_launch(where);
}

extern "C" int hn_painn_edge_bwd_dst_tiled(const hn_edge_params *p, const float *xh, const float *vec, const float *geom_b,
                                           const int32_t *bptr, const int32_t *meta, const int32_t *tile_rows,
                                           const int32_t *tile_mod, int32_t n_tiles, int32_t n_windows, const float *Wt,
                                           const float *bias, const float *offset, const float *g_dx, const float *g_dvec,
                                           float *g_geom_b, int64_t n_pad, void *stream) {
    const char *where = "hn_painn_edge_bwd_dst_tiled";
    if (int rc = validate_tiled(where, p, n_windows)) return rc;
    if (n_tiles <= 0) return 0;
    const size_t smem = kWarpsDst * (kAccDst * sizeof(float) + 512 + kStages * kStageBytes);
    const int ns = p->hidden / 64;
    const long long units = (long long)n_tiles * ns;
    dim3 grid((unsigned)((units + kWarpsDst - 1) / kWarpsDst));
    HN_DISPATCH_NS(p->hidden, {
        if (int rc = set_smem(edge_bwd_dst_tiled_kernel<NS>, smem, where)) return rc;
        edge_bwd_dst_tiled_kernel<NS><<<grid, 32 * kWarpsDst, smem, (cudaStream_t)stream>>>(
            *p, xh, vec, (const float4 *)geom_b, bptr, (const int4 *)meta, tile_rows, tile_mod, n_tiles, n_windows, Wt, bias,
            offset, g_dx, g_dvec, g_geom_b, (long long)n_pad);
    });
    return hn::check_launch(where);
}

extern "C" int hn_painn_edge_bwd_src_tiled(const hn_edge_params *p, const float *xh, const float *vec, const float *geom_s,
                                           const int32_t *bptr, const int32_t *meta, int32_t n_tiles, int32_t n_windows,
                                           const float *Wt, const float *bias, const float *offset, const float *g_dx,
                                           const float *g_dvec, float *grad_xh, float *grad_vec, void *stream) {
    const char *where = "hn_painn_edge_bwd_src_tiled";
    if (int rc = validate_tiled(where, p, n_windows)) return rc;
    if (n_tiles <= 0) return 0;
    const size_t smem = (size_t)kWarpsSrc * kWarpBytesSrc;
    const int ns = p->hidden / 64;
    const long long units = (long long)n_tiles * ns;
    dim3 grid((unsigned)((units + kWarpsSrc - 1) / kWarpsSrc));
    HN_DISPATCH_NS(p->hidden, {
        if (int rc = set_smem(edge_bwd_src_tiled_kernel<NS>, smem, where)) return rc;
        edge_bwd_src_tiled_kernel<NS><<<grid, 32 * kWarpsSrc, smem, (cudaStream_t)stream>>>(
            *p, xh, vec, (const float4 *)geom_s, bptr, (const int4 *)meta, n_tiles, n_windows, Wt, bias, offset, g_dx, g_dvec,
            grad_xh, grad_vec);
    });
    return hn::check_launch(where);
}
