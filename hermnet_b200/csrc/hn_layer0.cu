// First-layer edge pass of HVNet by basis aggregation, sm_100a.
//
// In the first layer x = Embedding[Z] and vec = 0 (HermNet/hermnet.py:123-124), so the projected source features of an edge
// only depend on the ELEMENT z of its source, and the message sum of rmnet.py:55-73 factorises: with
// phi_e = W (env * gauss(d_e)) + b (rmnet.py:55,168-193),
//     dx[i]      = sum_e xa[z_e] * phi_a(d_e)            = sum_z xa[z] * (W_a . Sa[i][z][:] + b_a * na[i][z])
//     dvec[i][c] = sum_e xc[z_e] * phi_c(d_e) * u_e[c]   = sum_z xc[z] * (W_c . Sc[i][c][z][:] + b_c * nc[i][c][z])
// where  Sa[i][z][k] = sum_{e -> i, z_e = z} env * gauss_k(d_e),  na = the edge count,  Sc / nc the same sums weighted by the
// unit vector.  The per-edge work shrinks from 3F = 384 channels to the 12-wide Gaussian band (x 4 weights), and the channel
// mixing becomes ONE dense GEMM per destination element over rows [Sa | na] (K' = n_elem * (K + 1), padded), which runs on the
// tensor cores (hn_gemm_tf32x3).  The backward pass mirrors it: g_S = g_d{x,vec} . Big^T (GEMM), then per edge
//     dL/dd = sum_k (g_Sa[z][k] + sum_c g_Sc[c][z][k] u_c) d/dd(env gauss_k),   dL/du_c = sum_k g_Sc[c][z][k] env gauss_k + g_nc[c][z].
//
// hn_layer0_basis_fwd / _bwd are the two per-edge kernels: one warp per destination row, the row's 4 x KP sums live in shared
// memory, lanes = (band position, weight pair); no atomics, deterministic.  Row layout: [z * K + k | n_elem * K + z | zero pad].
#include "hn_common.cuh"

namespace {

constexpr int kBand = 12, kBandLo = 5;      // Gaussian band floor(x)-5 .. floor(x)+6 (same band as hn_edge.cu / hn_edge_tc.cu)
constexpr int kWarps = 4;

struct L0Args {
    const int *rowptr, *elem, *row_mod;
    const float4 *geom;
    const unsigned char *live;     // NULL: every entry of the list is an edge
    const float *offset;
    long long n_rows;
    int K, nz, KP, env_p;
    float inv_rc, coeff;
};

__device__ __forceinline__ void envelope(float u, int p, float &env, float &denv) {
    const float a = -0.5f * (float)((p + 1) * (p + 2)), b = (float)(p * (p + 2)), c = -0.5f * (float)(p * (p + 1));
    float um = 1.f;
    for (int i = 0; i < p - 1; ++i) um *= u;
    const float u0 = um * u, u1 = u0 * u, u2 = u1 * u;
    env = 1.f + a * u0 + b * u1 + c * u2;
    denv = a * (float)p * um + b * (float)(p + 1) * u0 + c * (float)(p + 2) * u1;
}

// sum of val[0..15] over the 32 lanes; afterwards lane L holds the total of value index L & 15 (16 exchange shuffles)
__device__ __forceinline__ float transpose_reduce16(float (&val)[16], int lane) {
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const bool up = lane & w;
            const float send = up ? val[i] : val[i + w], keep = up ? val[i + w] : val[i];
            val[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    return val[0] + __shfl_xor_sync(0xffffffffu, val[0], 16);
}

// Shared memory: Gaussian centres [K] | per warp (forward): S[4 weights (1, ux, uy, uz)][KP + 8] + 64 scratch floats.
// Edges are read 32 at a time (one per lane, coalesced) and broadcast with shuffles; the per-edge body is branch-free (lanes
// without a band position update a private scratch word) and ends in __syncwarp(): the shared-memory updates of successive
// edges may hit the same word from different lanes.
template <bool BWD>
__global__ void __launch_bounds__(32 * kWarps) layer0_basis_kernel(L0Args A, float *__restrict__ Sa, float *__restrict__ Sc,
                                                                   const float *__restrict__ gSa, const float *__restrict__ gSc,
                                                                   float4 *__restrict__ g_geom) {
    extern __shared__ float4 smem4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KP = A.KP, K = A.K;
    float *offs = reinterpret_cast<float *>(smem4);
    const int koff = (K + 3) & ~3;
    for (int i = threadIdx.x; i < K; i += blockDim.x) offs[i] = __ldg(A.offset + i);
    __syncthreads();
    const long long row = (long long)blockIdx.x * kWarps + warp;
    if (row >= A.n_rows) return;
    // forward: the row's sums live in shared memory (read-modify-write per edge); backward: the g_S row is only read (48 words per
    // edge out of 4 KP) -- straight from global memory / L1, which leaves the occupancy to the registers
    // (row stride KP + 8 words: the two halves of the warp update rows 2 apart at the same column -- 2 (KP + 8) = 16 mod 32 puts
    // them on disjoint banks; with stride KP every update was a 2-way bank conflict and the L1 data pipe, at 78 %, set the time)
    const int KPs = KP + 8;
    float *S = offs + koff + (BWD ? 0 : (size_t)warp * (4 * KPs + 64));
    float *scratch = S + 4 * KPs + lane;
    const int e0 = __ldg(A.rowptr + row), e1 = __ldg(A.rowptr + row + 1);
    if (__ldg(A.row_mod + row) < 0) {
        if (BWD)
            for (int e = e0 + lane; e < e1; e += 32) g_geom[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const int n4 = KP >> 2;
    if (!BWD) {
        for (int i = lane; i < KPs + 16; i += 32) reinterpret_cast<float4 *>(S)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
    }
    const int kk = lane & 15, half = lane >> 4;
    const int cnt0 = A.nz * K;
    const float cl2 = A.coeff * 1.4426950408889634f;
    float *s0 = S + (2 * half) * KPs, *s1 = s0 + KPs;      // the two sums this lane owns: half 0 (1, ux), half 1 (uy, uz)
    const float *gc = BWD ? gSc + (size_t)row * 3 * KP : nullptr;
    const float *g0 = BWD ? (half == 0 ? gSa + (size_t)row * KP : gc + KP) : nullptr, *g1 = BWD ? (half == 0 ? gc : gc + 2 * KP) : nullptr;
    for (int eb = e0; eb < e1; eb += 32) {
        const int nb = min(32, e1 - eb);
        // lane l prepares edge eb + l: everything that only depends on the edge is computed once, then broadcast per edge
        float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);
        float ul = 2.f, envl = 0.f, de1l = 0.f, de0l = 0.f;
        int zkl = -1;                                      // z | kc << 8;  -1: no edge (beyond the row, or a dead entry of a Verlet-skin list)
        if (lane < nb) {
            gl = __ldg(A.geom + eb + lane);
            const int z = __ldg(A.elem + eb + lane);
            const bool dead = A.live != nullptr && A.live[eb + lane] == 0;
            ul = gl.w * A.inv_rc;
            float env, denv;
            envelope(fminf(ul, 1.f), A.env_p, env, denv);
            envl = ul < 1.f ? env : 0.f;
            de1l = envl * 2.f * A.coeff * A.inv_rc;        // d/dd (env gauss_k) = gauss_k * (de1 * diff + de0)
            de0l = ul < 1.f ? denv * A.inv_rc : 0.f;
            const int kc = min((int)(fminf(ul, 1.f) * (float)(K - 1)), K - 1);
            if (!dead) zkl = z | (kc << 8);
        }
        auto edge = [&](int j, float &w0, float &w1, float &val, float &dval, int &z, int &idx, bool &valid) {
            const float gx = __shfl_sync(0xffffffffu, gl.x, j), gy = __shfl_sync(0xffffffffu, gl.y, j),
                        gz = __shfl_sync(0xffffffffu, gl.z, j), u = __shfl_sync(0xffffffffu, ul, j),
                        env = __shfl_sync(0xffffffffu, envl, j);
            const int zk = __shfl_sync(0xffffffffu, zkl, j);
            z = zk < 0 ? -1 : (zk & 255);
            w0 = half == 0 ? 1.f : gy;
            w1 = half == 0 ? gx : gz;
            const int k = (zk >> 8) - kBandLo + kk;
            valid = zk >= 0 && u < 1.f && kk < kBand && k >= 0 && k < K;
            const int ks = min(max(k, 0), K - 1);
            const float diff = u - offs[ks];
            float gg;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(gg) : "f"(cl2 * diff * diff));
            gg = valid ? gg : 0.f;
            val = env * gg;
            dval = 0.f;
            if (BWD) dval = gg * fmaf(__shfl_sync(0xffffffffu, de1l, j), diff, __shfl_sync(0xffffffffu, de0l, j));
            idx = max(z, 0) * K + ks;
        };
        if (!BWD) {
            for (int j = 0; j < nb; ++j) {
                float w0, w1, val, dval;
                int z, idx;
                bool valid;
                edge(j, w0, w1, val, dval, z, idx, valid);
                // one update pair per lane: band position / edge count (lane 15 of each half) / private scratch word
                const bool cnt = kk == 15;
                const float v = cnt ? (z >= 0 ? 1.f : 0.f) : val;
                float *q0 = cnt ? s0 + cnt0 + max(z, 0) : (valid ? s0 + idx : scratch);
                float *q1 = cnt ? s1 + cnt0 + max(z, 0) : (valid ? s1 + idx : scratch + 32);
                *q0 = fmaf(v, w0, *q0);
                *q1 = fmaf(v, w1, *q1);
                __syncwarp();      // the next edge's band may touch these words from other lanes (racecheck-clean ordering)
            }
        } else {
            for (int j4 = 0; j4 < nb; j4 += 4) {
                float acc[16];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float w0, w1, val, dval;
                    int z, idx;
                    bool valid;
                    edge(j4 + jj, w0, w1, val, dval, z, idx, valid);      // (lanes >= nb hold zkl = -1: all zeros)
                    const float a0 = __ldg(g0 + idx), a1 = __ldg(g1 + idx);
                    acc[4 * jj + 0] = half == 0 ? a1 * val : 0.f;          // dL/dux
                    acc[4 * jj + 1] = half == 0 ? 0.f : a0 * val;          // dL/duy
                    acc[4 * jj + 2] = half == 0 ? 0.f : a1 * val;          // dL/duz
                    acc[4 * jj + 3] = fmaf(a0, w0, a1 * w1) * dval;        // dL/dd
                }
                const float tot = transpose_reduce16(acc, lane);            // lane & 15 = 4 * jj + component
                const int jj = kk >> 2, comp = kk & 3;
                const int zkj = __shfl_sync(0xffffffffu, zkl, (j4 + jj) & 31);
                const int zj = zkj < 0 ? -1 : (zkj & 255);
                if (lane < 16 && j4 + jj < nb) {
                    float out = tot;
                    if (comp < 3 && zj >= 0) out += __ldg(gc + comp * KP + cnt0 + zj);      // the count (filter-bias) terms
                    reinterpret_cast<float *>(g_geom + eb + j4 + jj)[comp] = zj >= 0 ? out : 0.f;
                }
            }
        }
    }
    if (!BWD) {
        __syncwarp();
        float4 *oa = reinterpret_cast<float4 *>(Sa + (size_t)row * KP);
        float4 *oc = reinterpret_cast<float4 *>(Sc + (size_t)row * 3 * KP);
        for (int i = lane; i < n4; i += 32) oa[i] = reinterpret_cast<const float4 *>(S)[i];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            for (int i = lane; i < n4; i += 32) oc[c * n4 + i] = reinterpret_cast<const float4 *>(S + (1 + c) * KPs)[i];
    }
}

int check(const char *where, const hn_edge_params *p, int32_t n_elem, int32_t kp) {
    HN_REQUIRE(p != nullptr, where, "null params");
    HN_REQUIRE(p->n_rows >= 0 && p->num_rbf >= 2 && p->env_p >= 1, where, "bad sizes");
    HN_REQUIRE(n_elem >= 1 && n_elem <= 255 && kp % 4 == 0 && (int64_t)kp >= (int64_t)n_elem * (p->num_rbf + 1), where,
               "row length must be a multiple of 4 and hold n_elem * (num_rbf + 1) sums");
    HN_REQUIRE((int64_t)kWarps * (4 * (kp + 8) + 64) * 4 + 4 * (int64_t)p->num_rbf <= 200 * 1024, where, "row too long for shared memory");
    return 0;
}

template <bool BWD>
int launch(const char *where, const hn_edge_params *p, const int32_t *rowptr, const int32_t *elem, const int32_t *row_mod,
           const float *geom, const uint8_t *live, const float *offset, int32_t n_elem, int32_t kp, float *Sa, float *Sc,
           const float *gSa, const float *gSc, float *g_geom, void *stream) {
    if (int rc = check(where, p, n_elem, kp)) return rc;
    if (p->n_rows == 0) return 0;
    L0Args a;
    a.rowptr = rowptr; a.elem = elem; a.row_mod = row_mod; a.geom = reinterpret_cast<const float4 *>(geom);
    a.live = (p->flags & 1) ? live : nullptr;
    a.offset = offset; a.n_rows = p->n_rows; a.K = p->num_rbf; a.nz = n_elem; a.KP = kp; a.env_p = p->env_p;
    a.inv_rc = p->inv_rc; a.coeff = p->coeff;
    HN_REQUIRE(!(p->flags & 1) || live != nullptr, where, "a Verlet-skin list (flags bit 0) needs the live mask");
    const size_t smem = ((BWD ? 0 : (size_t)kWarps * (4 * (kp + 8) + 64)) + ((p->num_rbf + 3) & ~3)) * sizeof(float);
    HN_CUDA(cudaFuncSetAttribute(layer0_basis_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), where);
    const unsigned blocks = (unsigned)((p->n_rows + kWarps - 1) / kWarps);
    layer0_basis_kernel<BWD><<<blocks, 32 * kWarps, smem, (cudaStream_t)stream>>>(a, Sa, Sc, gSa, gSc, reinterpret_cast<float4 *>(g_geom));
    return hn::check_launch(where);
}

}  // namespace

extern "C" int hn_layer0_basis_fwd(const hn_edge_params *p, const int32_t *rowptr, const int32_t *elem, const int32_t *row_mod,
                                   const float *geom, const uint8_t *live, const float *offset, int32_t n_elem, int32_t kp,
                                   float *Sa, float *Sc, void *stream) {
    return launch<false>("hn_layer0_basis_fwd", p, rowptr, elem, row_mod, geom, live, offset, n_elem, kp, Sa, Sc, nullptr, nullptr,
                         nullptr, stream);
}

extern "C" int hn_layer0_basis_bwd(const hn_edge_params *p, const int32_t *rowptr, const int32_t *elem, const int32_t *row_mod,
                                   const float *geom, const uint8_t *live, const float *offset, int32_t n_elem, int32_t kp,
                                   const float *g_Sa, const float *g_Sc, float *g_geom, void *stream) {
    return launch<true>("hn_layer0_basis_bwd", p, rowptr, elem, row_mod, geom, live, offset, n_elem, kp, nullptr, nullptr, g_Sa, g_Sc,
                        g_geom, stream);
}
