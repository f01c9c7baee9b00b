// Halo exchange of the domain-decomposed path (north_star "ghost-atom halo exchange ... over NCCL/NVLink"):
//   hn_halo_pack    gathers the feature rows  [x (F) | vec (3F)]  of the atoms on a send list and stores every row
//                   STRAIGHT INTO THE DESTINATION RANK'S landing buffer -- `dst_base[peer]` are device pointers, the peers'
//                   buffers mapped into this process (NVLink peer memory; symmetric buffers exchanged once per
//                   decomposition), or, on the fallback path, offsets into one local send buffer for an NCCL all-to-all;
//   hn_halo_unpack  moves the landed rows into the ghost rows of x / vec.
// The backward pass uses the same two kernels with the roles of the lists swapped (ghost-row gradients are packed into the
// OWNERS' landing buffers -- the reverse force accumulation -- and reduced there with hn_segment_sum).
// The reference has no counterpart: its only multi-GPU path is DDP (example/dist_train.py).
#include "hn_common.cuh"

namespace {

// one warp per row: W4 = row width in float4 (x: F/4, vec: 3F/4, contiguous in the landing buffer)
__global__ void halo_pack_kernel(const float4 *__restrict__ x, const float4 *__restrict__ vec, const int32_t *__restrict__ src_idx,
                                 const int32_t *__restrict__ row_peer, const int32_t *__restrict__ row_slot,
                                 const unsigned long long *__restrict__ dst_base, long long n_rows, int F4) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const long long s = src_idx[w];
    float4 *dst = reinterpret_cast<float4 *>(dst_base[row_peer[w]]) + (long long)row_slot[w] * (4 * F4);
    const float4 *xs = x + s * F4, *vs = vec + s * (3 * F4);
    for (int i = lane; i < F4; i += 32) dst[i] = __ldg(xs + i);
    for (int i = lane; i < 3 * F4; i += 32) dst[F4 + i] = __ldg(vs + i);
}

__global__ void halo_unpack_kernel(const float4 *__restrict__ buf, const int32_t *__restrict__ dst_idx, long long n_rows, int F4,
                                   float4 *__restrict__ x, float4 *__restrict__ vec) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const long long d = dst_idx[w];
    const float4 *src = buf + w * (4 * F4);
    float4 *xd = x + d * F4, *vd = vec + d * (3 * F4);
    for (int i = lane; i < F4; i += 32) xd[i] = src[i];
    for (int i = lane; i < 3 * F4; i += 32) vd[i] = src[F4 + i];
}

}  // namespace

extern "C" int hn_halo_pack(const float *x, const float *vec, const int32_t *src_idx, const int32_t *row_peer,
                            const int32_t *row_slot, const uint64_t *dst_base, int64_t n_rows, int32_t hidden, void *stream) {
    const char *where = "hn_halo_pack";
    if (n_rows <= 0) return 0;
    HN_REQUIRE(hidden >= 4 && hidden % 4 == 0, where, "hidden_channels must be a multiple of 4");
    HN_REQUIRE((((uintptr_t)x | (uintptr_t)vec) & 15) == 0, where, "x / vec must be 16-byte aligned");
    const long long blocks = (n_rows * 32 + 255) / 256;
    halo_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)x, (const float4 *)vec, src_idx, row_peer,
                                                                         row_slot, (const unsigned long long *)dst_base, n_rows,
                                                                         hidden / 4);
    return hn::check_launch(where);
}

extern "C" int hn_halo_unpack(const float *buf, const int32_t *dst_idx, int64_t n_rows, int32_t hidden, float *x, float *vec,
                              void *stream) {
    const char *where = "hn_halo_unpack";
    if (n_rows <= 0) return 0;
    HN_REQUIRE(hidden >= 4 && hidden % 4 == 0, where, "hidden_channels must be a multiple of 4");
    HN_REQUIRE((((uintptr_t)x | (uintptr_t)vec | (uintptr_t)buf) & 15) == 0, where, "buffers must be 16-byte aligned");
    const long long blocks = (n_rows * 32 + 255) / 256;
    halo_unpack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)buf, dst_idx, n_rows, hidden / 4,
                                                                           (float4 *)x, (float4 *)vec);
    return hn::check_launch(where);
}
