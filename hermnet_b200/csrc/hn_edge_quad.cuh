// Internal interface of the quad-tile edge kernels (hn_edge_quad.cu); dispatched from the C-ABI entry points in hn_edge.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hermnet_b200.h"

namespace hn {
namespace quad {

bool supported(const hn_edge_params *p);   // F % 64 == 0
int bwd_dst_slices(int hidden);            // planes of g_geom written by bwd_dst

int fwd(const hn_edge_params *p, const float *xh, const float *vec, const float *geom, const int32_t *rowptr,
        const int32_t *col, const int32_t *row_mod, const int64_t *row_xoff, const float *Wt, const float *bias,
        const float *offset, float *dx, float *dvec, cudaStream_t stream);
int bwd_dst(const hn_edge_params *p, const float *xh, const float *vec, const float *geom, const int32_t *rowptr,
            const int32_t *col, const int32_t *row_mod, const int64_t *row_xoff, const float *Wt, const float *bias,
            const float *offset, const float *g_dx, const float *g_dvec, float *g_geom, int64_t n_edges, cudaStream_t stream);

int bwd_src(const hn_edge_params *p, const float *xh, const float *vec, const float *geom, const int32_t *t_rowptr,
            const int32_t *t_eid, const int32_t *edge_row, const int32_t *row_mod, const int64_t *row_xoff, const float *Wt,
            const float *bias, const float *offset, const float *g_dx, const float *g_dvec, float *grad_xh, float *grad_vec,
            cudaStream_t stream);

}  // namespace quad
}  // namespace hn
