/* hermnet_b200 -- C ABI of the B200-native HermNet message-passing hot path.
 *
 * The reference (thu-wangz17/HermNet) has no FFI: its hot path is a chain of implicit PyTorch /
 * PyG / torch_scatter / torch_cluster / ASE library calls.  Each entry point below names the
 * reference call site(s) it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  Every pointer is a DEVICE pointer on the
 *     current CUDA device unless stated otherwise; the caller owns every buffer (outputs, scratch).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - return value: 0 = ok, non-zero = error; `hn_last_error()` gives a thread-local message.
 *   - variable-size outputs use two phases (count -> caller scans/allocates -> fill).
 *   - indices are int32 (E < 2^31); triplet offsets are int64.
 *   - float data is fp32; the neighbour test runs in fp64 (see hn_radius_graph_*).
 *
 * Graph layout ("row CSR")
 *   A *row* is one segment of the segmented reduction: (destination atom [, module slot]).
 *   rows_per_atom is constant: row r belongs to atom r / rows_per_atom.
 *     rowptr[R+1], col[E] (source atom of each row-edge), shift[E][4] (int8 Sx,Sy,Sz,0),
 *     row_mod[R] (module / weight-set id of the row, -1 = inactive row).
 *   The transposed view (grouped by source atom) is t_rowptr[N+1], t_eid[E] (row-edge ids),
 *   edge_row[E] (row of each row-edge).
 */
#ifndef HERMNET_B200_H
#define HERMNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HN_ABI_VERSION 1

int hn_abi_version(void);
const char *hn_last_error(void);
/* number of SMs of the current device (grid sizing), <0 on error */
int hn_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Radius graph: cell-list neighbour search producing a centre-sorted CSR.
 * Replaces HermNet/data.py:14-24 `neighbor_search` = ase.neighborlist.primitive_neighbor_list
 * (periodic, data.py:19-21) / torch_cluster radius_graph (non-periodic, data.py:16).
 *
 * Row c*n_groups+g lists every (j, S) with group[j]==g and
 *     || pos_j - pos_c + S.cell || < rc ,  not (j==c and S==0)
 * evaluated as  sqrt(sum((double)(float)(pos_j-pos_c) + S.cell)^2) < rc  in fp64, un-fused
 * (the ASE arithmetic).  cell==NULL: non-periodic, fp32 test d2 < rc*rc, S==0, and at most
 * max_neighbors (>0) neighbours per centre, the smallest indices (torch_cluster CUDA semantics).
 * Positions need not lie inside the cell.  Atoms of a graph are contiguous: graph_ptr[B+1].
 * `workspace` (hn_radius_graph_workspace_bytes) carries the cell list from count to fill.
 * Entry order inside a row is deterministic (cell-traversal order).
 * ------------------------------------------------------------------------------------------- */
int64_t hn_radius_graph_workspace_bytes(int64_t n_atoms, int32_t n_graphs);
int hn_radius_graph_count(const float *pos, int64_t n_atoms, const float *cell /*[B,3,3] or NULL*/,
                          const int32_t *graph_ptr, int32_t n_graphs, double rc,
                          const int32_t *group /*[N] or NULL*/, int32_t n_groups, int32_t max_neighbors,
                          int32_t *counts /*[N*n_groups]*/, void *workspace, int64_t workspace_bytes,
                          void *stream);
int hn_radius_graph_fill(const float *pos, int64_t n_atoms, const float *cell, const int32_t *graph_ptr,
                         int32_t n_graphs, double rc, const int32_t *group, int32_t n_groups,
                         int32_t max_neighbors, const int32_t *rowptr /*[N*n_groups+1]*/,
                         int32_t *col /*[E]*/, int8_t *shift /*[E,4]*/, void *workspace,
                         int64_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Stable counting sort of edge ids by an integer key: order[] lists ids grouped by key ascending,
 * original order inside a key; rowptr[n_keys+1].  Used for (a) COO -> row CSR of a user-supplied
 * edge_index (replaces the per-atom `torch.where` scans of HermNet/utils.py:11-24 `in_subgraph`),
 * (b) the transposed (source-major) view needed by the backward pass.
 * ------------------------------------------------------------------------------------------- */
int64_t hn_sort_by_key_workspace_bytes(int64_t n, int32_t n_keys);
int hn_sort_by_key(const int32_t *keys, int64_t n, int32_t n_keys, int32_t *rowptr, int32_t *order,
                   void *workspace, int64_t workspace_bytes, void *stream);
/* edge_row[e] = r for rowptr[r] <= e < rowptr[r+1] */
int hn_expand_rowptr(const int32_t *rowptr, int32_t n_rows, int32_t *edge_row, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Canonical ordered triplets (j,i,k) over row-edge ids (HTNet -- not in the
 * reference, HermNet/hermnet.py:155-157 raises): all (e1,e2), e1 != e2, in the same row, sorted
 * by (row, e1, e2).  Typed variant: src_type[col[e1]]==type_a and src_type[col[e2]]==type_c
 * (pass src_type=NULL for untyped).
 * ------------------------------------------------------------------------------------------- */
int hn_triplets_count(const int32_t *rowptr, int32_t n_rows, const int32_t *col, const int32_t *src_type,
                      int32_t type_a, int32_t type_c, int64_t *counts /*[R]*/, void *stream);
int hn_triplets_fill(const int32_t *rowptr, int32_t n_rows, const int32_t *col, const int32_t *src_type,
                     int32_t type_a, int32_t type_c, const int64_t *trip_ptr /*[R+1]*/,
                     int32_t *e1 /*[T]*/, int32_t *e2 /*[T]*/, void *stream);
/* dots[r][f] = sum over triplets (e1,e2) of row r of sum_k m_vec[e1][k][f]*m_vec[e2][k][f] */
int hn_triplet_dots(const float *m_vec /*[E,3,F]*/, int32_t F, const int64_t *trip_ptr, const int32_t *e1,
                    const int32_t *e2, int32_t n_rows, float *dots /*[R,F]*/, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Edge geometry.  Replaces HVNet.with_edge, HermNet/hermnet.py:133-152:
 *   D = pos[col[e]] - pos[atom(row(e))] + sign * S_e . cell[graph[col[e]]]   (sign=+1: reference
 *   quirk F5, -1: physical);  d = |D|, d <= 1e-6 -> 1e-6;  u = D/d.   geom[e] = (ux,uy,uz,d).
 * Backward: g_geom[n_parts][E] = (dL/dux, dL/duy, dL/duz, dL/dd), summed over parts, gives
 *   grad_pos[N,3] (segmented sums over the row CSR and its transpose -- no atomics) and, if
 *   cellw != NULL, the per-atom virial partial cellw[i][a][b] = sign * sum_{e in rows(i)} S_e[a]*gD_e[b].
 * ------------------------------------------------------------------------------------------- */
int hn_edge_geom_fwd(const float *pos, const float *cell, const int32_t *atom_graph, const int32_t *edge_row,
                     int32_t rows_per_atom, const int32_t *col, const int8_t *shift, float sign, int64_t n_edges,
                     float *geom /*[E,4]*/, void *stream);
int hn_edge_geom_bwd(const float *geom, const float *g_geom, int32_t n_parts, const int8_t *shift,
                     const int32_t *rowptr, int32_t rows_per_atom, const int32_t *t_rowptr, const int32_t *t_eid,
                     float sign, int64_t n_atoms, int64_t n_edges, float *grad_pos /*[N,3]*/,
                     float *cellw /*[N,9] or NULL*/, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused PaiNN edge kernel.  Replaces, per layer and per sub-network, HermNet/rmnet.py:55-73
 * (rbf_proj, propagate/message, scatter aggregate), rmnet.py:168-193 (Gaussian RBF x polynomial
 * envelope) and the sub-graph regrouping of HermNet/utils.py:11-24 / hermnet.py:51-61:
 *   phi_e = W[m] . (env(d/rc) * gauss_k(d/rc)) + b[m]                       (3F)
 *   dx[r]      = sum_e  xh[m][s][0:F]  * phi_e[0:F]
 *   dvec[r][k] = sum_e (vec[s][k] * xh[m][s][F:2F]*phi_e[F:2F]/sqrt(3) + xh[m][s][2F:3F]*phi_e[2F:3F]*u_e[k]) / sqrt(F)
 * with m = row_mod[r], s = col[e]; segmented reduction per row, warp shuffles, no global atomics.
 * xh is ONE flat fp32 buffer holding the projected source features of every sub-network compactly;
 * row r reads source s at row (row_xoff[r] + s) of that [rows][3F] buffer (int64 row offsets; for a dense
 * [M][N][3F] layout row_xoff[r] = m*N).  Wt is rbf_proj.weight transposed: [M][K][3F].  gauss_k uses the `offset` buffer (K values,
 * linspace(0,1,K)) and coeff = -0.5/(offset[1]-offset[0])^2; F % 32 == 0.
 * vec may be NULL in fwd and bwd_dst: vec == 0 identically (the first layer, hermnet.py:124), which removes the vec
 * gathers and the F:2F part of the filter from those two passes.
 * bwd_dst (row-major pass): g_geom[n_slices][E] = per-edge (dL/du, dL/dd).
 * bwd_src (source-major pass over the transpose): grad_xh (same flat layout as xh; must be zero-filled
 *   by the caller) and grad_vec[N][3][F].
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_atoms;      /* N: rows of vec / of each xh[m] */
    int32_t n_rows;       /* R */
    int32_t n_modules;    /* M */
    int32_t hidden;       /* F */
    int32_t num_rbf;      /* K */
    int32_t env_p;        /* polynomial envelope exponent */
    float inv_rc;         /* 1/cutoff */
    float coeff;          /* Gaussian coefficient */
    int32_t variant;      /* kernel family of hn_painn_edge_{fwd,bwd_dst,bwd_src}: 0 = auto (quad-tile kernels for F % 64 == 0,
                             else row-per-warp), 1 = row-per-warp only (A/B measurements).  Per call: no library state. */
    int32_t flags;        /* bit 0: the edge list is a Verlet-skin SUPERSET (built with rc + skin): the entries marked dead
                             by hn_tc_tile_windows(live) must contribute nothing -- not even the filter bias (rmnet.py:55:
                             the reference has no such edge).  Honoured by the hn_tc_edge_* kernels; the other families
                             reject it. */
} hn_edge_params;

/* Planes of g_geom that hn_painn_edge_bwd_dst writes for this F and kernel family (the caller sums them); 0 = unsupported F. */
int32_t hn_painn_edge_num_slices(int32_t hidden, int32_t variant);
int hn_painn_edge_fwd(const hn_edge_params *p, const float *xh, const float *vec, const float *geom,
                      const int32_t *rowptr, const int32_t *col, const int32_t *row_mod, const int64_t *row_xoff,
                      const float *Wt, const float *bias, const float *offset, float *dx /*[R,F]*/,
                      float *dvec /*[R,3,F]*/, void *stream);
int hn_painn_edge_bwd_dst(const hn_edge_params *p, const float *xh, const float *vec, const float *geom,
                          const int32_t *rowptr, const int32_t *col, const int32_t *row_mod, const int64_t *row_xoff,
                          const float *Wt, const float *bias, const float *offset, const float *g_dx,
                          const float *g_dvec, float *g_geom /*[n_slices,E,4]*/, int64_t n_edges, void *stream);
int hn_painn_edge_bwd_src(const hn_edge_params *p, const float *xh, const float *vec, const float *geom,
                          const int32_t *t_rowptr, const int32_t *t_eid, const int32_t *edge_row,
                          const int32_t *row_mod, const int64_t *row_xoff, const float *Wt, const float *bias,
                          const float *offset, const float *g_dx, const float *g_dvec, float *grad_xh /*zeroed*/,
                          float *grad_vec /*[N,3,F]*/, void *stream);

/* Filter-weight gradient of the fused path (autograd of nn.Linear rbf_proj, rmnet.py:45,55):
 *   gW_part[chunk][m][k][c] / gb_part[chunk][m][c] hold per-row-chunk partial sums of
 *   gW[m][k][c] = sum_e env*gauss_k(d_e) * dL/dphi_e[c],  gb[m][c] = sum_e dL/dphi_e[c];
 *   the caller adds the n_chunks partials in fixed order (deterministic). */
int hn_painn_edge_bwd_w(const hn_edge_params *p, const float *xh, const float *vec, const float *geom,
                        const int32_t *rowptr, const int32_t *col, const int32_t *row_mod, const int64_t *row_xoff,
                        const float *offset, const float *g_dx, const float *g_dvec,
                        float *gW_part /*[n_chunks,M,K,3F]*/, float *gb_part /*[n_chunks,M,3F]*/, int32_t n_chunks,
                        void *stream);

/* ---------------------------------------------------------------------------------------------
 * First-layer edge pass of HVNet by basis aggregation.  In the first layer x = Embedding[Z], vec = 0
 * (HermNet/hermnet.py:123-124): the projected source features depend only on the source ELEMENT z, so the message sum of
 * rmnet.py:55-73 is linear in per-(destination, element) sums of the radial basis:
 *   Sa[r][z*K + k]        = sum_{e -> r, elem[e] = z} env * gauss_k(d_e)          Sa[r][n_elem*K + z]    = edge count
 *   Sc[r][c][z*K + k]     = sum ... * u_e[c]                                      Sc[r][c][n_elem*K + z] = sum u_e[c]
 * (rows of `kp` floats, kp % 4 == 0, kp >= n_elem * (K + 1); the pad is written as zeros), and
 *   dx[r] = [Sa[r]] . BigA[m]^T,  dvec[r][c] = [Sc[r][c]] . BigC[m]^T   with  BigA[m][f][z*K+k] = xa[m][z][f] * W_a[m][k][f],
 *   BigA[m][f][n_elem*K+z] = xa[m][z][f] * b_a[m][f]  (BigC likewise with xc / sqrt(F)) -- dense GEMMs (hn_gemm_tf32x3).
 * `elem[e]` = element index of the source of row-edge e; rows with row_mod < 0 are skipped (fwd: their S rows are not
 * written; bwd: zeros in g_geom).  bwd: g_geom[e] = (dL/du, dL/dd) from g_Sa / g_Sc (same layout as Sa / Sc).
 * p: n_rows, num_rbf, env_p, inv_rc, coeff, flags (bit 0: `live` [E] uint8 marks the entries of a Verlet-skin list that are
 * edges right now; else live may be NULL).
 * ------------------------------------------------------------------------------------------- */
int hn_layer0_basis_fwd(const hn_edge_params *p, const int32_t *rowptr, const int32_t *elem, const int32_t *row_mod,
                        const float *geom, const uint8_t *live, const float *offset, int32_t n_elem, int32_t kp,
                        float *Sa /*[R,kp]*/, float *Sc /*[R,3,kp]*/, void *stream);
int hn_layer0_basis_bwd(const hn_edge_params *p, const int32_t *rowptr, const int32_t *elem, const int32_t *row_mod,
                        const float *geom, const uint8_t *live, const float *offset, int32_t n_elem, int32_t kp,
                        const float *g_Sa, const float *g_Sc, float *g_geom /*[E,4]*/, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core PaiNN edge kernels (tcgen05 / TMEM / TMA; hidden_channels == 128, num_rbf <= 256).
 * Same arithmetic and reference lines as hn_painn_edge_* above; the filter projection
 * rbf_proj (HermNet/rmnet.py:45,55) runs as fp16-split (hi + lo) tensor-core tiles with fp32
 * accumulation: PHI^T[3F x 64 edges] = W^T[3F x 32-wide basis window] . BASIS^T.
 *
 * Tile plan (built once per graph, replaces the per-layer regrouping of HermNet/utils.py:11-24):
 *   a *block* is up to hn_tc_block_rows(0|1) rows of ONE sub-network (dst-major: row0 + l*stride) or
 *   consecutive source atoms (src-major); a *tile* is up to hn_tc_tile_edges() edges of a block
 *   (src-major: of one (block, sub-network) group) whose Gaussian bands fit one 32-wide window.
 *     blk_info[n_blocks][4]  dst-major (row0, row stride, n_rows, module | -1), src-major (first atom, 1, n, -1)
 *     blk_tile[n_blocks+1]   tile range of each block
 *     tile_info[n_tiles][4]  (first edge record, count, cumulative ends of the (local % G) groups packed 8 bits each, module)
 *     tile_win[n_tiles][2]   (k0, n_chunks): written by hn_tc_tile_windows for the CURRENT geometry, together with
 *     tile_geom[E][4]        the geometry (ux,uy,uz,d) of every record in record order (the kernels read THIS, not geom)
 *     erec[E][4]             dst-major (xh row, source atom, row_local, edge id),
 *                            src-major (destination row, xh row of the source, source_local, edge id)
 *   Build: kc = hn_tc_basis_index(geom); order = edges sorted by (group, kc) (hn_sort_by_key twice);
 *   hn_tc_plan_count -> caller scans -> hn_tc_plan_fill (tile_start) -> hn_tc_plan_finalize, which
 *   sorts every tile by (local % G, local) and writes erec / tile_info from rec[E][4] (indexed by
 *   edge id, field 2 = local index) and tile_mod[n_tiles].
 * Weights: hn_tc_split_weights turns Wt [M][K][3F] into fp16 hi and lo planes [2][M*3F][K32],
 *   K32 = num_rbf rounded up to 32, scaled by a per-module power of two; wscale[m] = 1/scale.
 * Outputs as in hn_painn_edge_*: fwd dx/dvec (every row of every block is written), bwd_dst one
 *   g_geom plane [E][4] (only edges in tiles are written), bwd_src grad_xh (zero-filled by the
 *   caller) and grad_vec (every atom of every block is written).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_blocks;
    int32_t n_tiles;
    const int32_t *blk_info;
    const int32_t *blk_tile;
    const int32_t *blk_xoff;   /* dst-major [n_blocks]: xh row of source s for the block's rows = blk_xoff[b] + erec[.][1] */
    const int32_t *tile_info;
    int32_t *tile_win;
    const int32_t *erec;
    float *tile_geom;          /* [E][4]: geometry of every record, written by hn_tc_tile_windows */
    const float *zero_row;     /* >= 3F zeros (required when hn_edge_params.flags bit 0 is set, else may be NULL) */
    const int32_t *blk_order;  /* [n_blocks] processing order of the blocks (a permutation; NULL = identity).  Results do not
                                  depend on it; interleaving the element-type slices spatially keeps the rows that several
                                  sub-networks gather (vec, g_dx / g_dvec) in L2 between their uses. */
} hn_tc_plan;

int32_t hn_tc_supported(int32_t hidden, int32_t num_rbf);
int32_t hn_tc_block_rows(int32_t src_major);   /* rows (dst-major) / source atoms (src-major) per block */
int32_t hn_tc_tile_edges(void);
int32_t hn_tc_groups(void);          /* epilogue groups G: tiles are sorted by (local % G, local) */
int64_t hn_tc_split_weights_elems(int32_t n_modules, int32_t hidden, int32_t num_rbf);   /* fp16 elements of wsplit */
int hn_tc_split_weights(const float *Wt, int32_t n_modules, int32_t num_rbf, int32_t hidden, void *wsplit /*fp16*/,
                        float *wscale /*[M]*/, void *stream);
int hn_tc_basis_index(const float *geom, int64_t n_edges, float inv_rc, int32_t num_rbf, int32_t *kc /*[E]*/, void *stream);
/* Per-edge inputs of a plan in one pass: kc[e] = basis index of the edge (as hn_tc_basis_index), rec[e][4] = its record
 * (dst-major: xh row, source, row-local index = atom_local[row / rows_per_atom], edge id; src-major: destination row, xh row,
 * source % src_block, edge id), sub[e] = sort sub-key (dst-major: 0, src-major: sub-network of the row; -1 = inactive row). */
int hn_tc_plan_records(int32_t src_major, const int32_t *edge_row, const int32_t *col, const int64_t *row_xoff, const int32_t *row_mod,
                       const int32_t *atom_local, int32_t rows_per_atom, int32_t src_block, const float *geom, float inv_rc,
                       int32_t num_rbf, int64_t n_edges, int32_t *rec, int32_t *kc, int32_t *sub, void *stream);
/* Segment-local sort of the plan builder: segment s = entries in_ptr[s] .. in_ptr[s+1] of the edge list `ids` (NULL: identity)
 * -- the CSR entries of a destination block or the transposed-CSR entries of a source block; entries are ordered by (sub, kc)
 * inside their segment (sub: NULL = 0; -1 drops the entry), stable and deterministic.  Count pass: counts[s * n_sub + sub]
 * (order_out NULL); fill pass: order_out[out_base[s] + position] = edge id (counts NULL).  n_sub * num_rbf <= 8192. */
int hn_tc_plan_sort(const int32_t *in_ptr, const int32_t *ids, const int32_t *kc, const int32_t *sub, int32_t n_seg, int32_t n_sub,
                    int32_t num_rbf, int32_t *counts, const int32_t *out_base, int32_t *order_out, void *stream);
int hn_tc_plan_count(const int32_t *order, const int32_t *kc, const int32_t *grp_ptr, int32_t n_groups, int32_t num_rbf,
                     int32_t window /* 32, 64, ...: basis-index width of a tile (k-chunks per tile = window / 32) */,
                     int32_t *counts, void *stream);
int hn_tc_plan_fill(const int32_t *order, const int32_t *kc, const int32_t *grp_ptr, int32_t n_groups, int32_t num_rbf,
                    int32_t window, const int32_t *grp_tile, int32_t *tile_start, void *stream);
int hn_tc_plan_finalize(const int32_t *order, const int32_t *tile_start, int32_t n_tiles, int64_t n_edges,
                        const int32_t *rec /*[E][4]*/, const int32_t *tile_mod, int32_t *erec, int32_t *tile_info, void *stream);
/* live: NULL, or uint8 [E] -- 0 marks an entry of a Verlet-skin superset list that is not an edge now (stored as -d in
 * tile_geom; with hn_edge_params.flags bit 0 the edge kernels make it contribute nothing) */
int hn_tc_tile_windows(const hn_tc_plan *plan, const float *geom, const uint8_t *live, float inv_rc, int32_t num_rbf, void *stream);
int hn_tc_edge_fwd(const hn_edge_params *p, const hn_tc_plan *plan, const float *xh, const float *vec /*NULL: vec == 0*/,
                   const float *geom, const void *wsplit, const float *wscale, const float *bias, const float *offset,
                   float *dx, float *dvec, float *dbg /*NULL, or 2*E*3F + 14336 floats: phi | raw accumulators | first operand stage*/,
                   int64_t dbg_edges /*E of the debug layout*/, void *stream);
int hn_tc_edge_bwd_dst(const hn_edge_params *p, const hn_tc_plan *plan, const float *xh, const float *vec, const float *geom,
                       const void *wsplit, const float *wscale, const float *bias, const float *offset, const float *g_dx,
                       const float *g_dvec, float *g_geom /*[E][4]*/, void *stream);
int hn_tc_edge_bwd_src(const hn_edge_params *p, const hn_tc_plan *plan /*src-major*/, const float *xh, const float *vec,
                       const float *geom, const void *wsplit, const float *wscale, const float *bias, const float *offset,
                       const float *g_dx, const float *g_dvec, float *grad_xh /*zeroed*/, float *grad_vec, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Row gather / segmented sum -- mutually adjoint linear primitives used by the differentiable
 * (double-backward, training) formulation.  Replace PyG's index_select gathers (rmnet.py:58) and
 * torch_scatter.scatter's atomicAdd (rmnet.py:71-72, hermnet.py:130) with deterministic kernels.
 *   gather:       out[e][:] = X[idx[e]][:]
 *   segment_sum:  out[r][:] = sum_{q in [rowptr[r],rowptr[r+1])} Y[perm ? perm[q] : q][:]
 * Also used for halo pack (gather) / reverse halo accumulation (segment_sum) in the domain-
 * decomposed path.
 * ------------------------------------------------------------------------------------------- */
int hn_gather_rows(const float *X, const int32_t *idx, int64_t n_out, int32_t C, float *out, void *stream);
int64_t hn_segment_sum_workspace_bytes(int32_t n_rows, int32_t C);   /* 0 when no workspace is needed */
int hn_segment_sum(const float *Y, const int32_t *rowptr, const int32_t *perm, int32_t n_rows, int32_t C,
                   float *out, void *workspace, int64_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Row normalisation of the node features: the nn.LayerNorm of rmnet.py:39,52 without its affine part (the caller folds
 * weight / bias into the Linear that follows).  xhat[i] = (x[i] - mean[i]) * rstd[i], rstd = 1/sqrt(var + eps), biased
 * variance over the `hidden` channels; hidden % 32 == 0, <= 512.  Backward: g_x from g_xhat and the saved x, mean, rstd.
 * ------------------------------------------------------------------------------------------- */
int hn_layernorm_fwd(const float *x, int64_t n, int32_t hidden, float eps, float *xhat, float *mean, float *rstd, void *stream);
int hn_layernorm_bwd(const float *g_xhat, const float *x, const float *mean, const float *rstd, int64_t n, int32_t hidden,
                     float *g_x, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Readout MLP (hermnet.py:112-116,129): e_atom[i] = W2 . ssilu(W1 x[i] + b1) + b2, W1 [F/2, F], W2 [F/2], plain fp32 FMAs
 * (the atomic energies cancel strongly in the sum; this layer does not use the 3xTF32 GEMM).  bwd: g_x[i] = dE/dx[i] given
 * g_e[i] = dL/de_atom[i]; the hidden activations are recomputed.  F = 64 or 128 (W1 lives in shared memory); b2: device
 * pointer to the scalar bias (no host read-back: the call can be captured in a CUDA graph).
 * ------------------------------------------------------------------------------------------- */
int hn_readout_fwd(const float *x, const float *W1, const float *b1, const float *W2, const float *b2, int64_t n, int32_t hidden,
                   float *e_atom, void *stream);
int hn_readout_bwd(const float *x, const float *W1, const float *b1, const float *W2, const float *b2, const float *g_e, int64_t n,
                   int32_t hidden, float *g_x, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Halo exchange of the domain-decomposed path (no counterpart in the reference, whose only
 * multi-GPU path is DDP, example/dist_train.py).  A landing-buffer row is [x (F) | vec (3F)] floats.
 *   hn_halo_pack:   row i = features of atom src_idx[i]; stored at
 *                   ((float *)dst_base[row_peer[i]]) + row_slot[i] * 4F -- dst_base[] holds device addresses: the PEERS'
 *                   landing buffers mapped into this process (NVLink peer memory), or slices of one local send buffer
 *                   (NCCL all-to-all fallback).
 *   hn_halo_unpack: row i of buf -> x[dst_idx[i]], vec[dst_idx[i]] (the ghost rows).
 * The backward pass calls the same pair with the lists swapped (ghost-row gradients into the owners' buffers, then
 * hn_segment_sum): the reverse force accumulation.
 * ------------------------------------------------------------------------------------------- */
int hn_halo_pack(const float *x, const float *vec, const int32_t *src_idx, const int32_t *row_peer, const int32_t *row_slot,
                 const uint64_t *dst_base, int64_t n_rows, int32_t hidden, void *stream);
int hn_halo_unpack(const float *buf, const int32_t *dst_idx, int64_t n_rows, int32_t hidden, float *x, float *vec, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Node-side dense layers on the tcgen05 tensor cores (3xTF32 split, fp32-class accuracy):
 *   C[M,N] (row pitch ldc) = A[M,K] (row pitch lda) . W[N,K]^T + bias[N]   (bias may be NULL)
 * Replaces the cuBLAS fp32 GEMMs behind nn.Linear in HermNet/rmnet.py:40-49 (x_proj), :84-89
 * (vec_proj, xvec_proj) and hermnet.py:112-116 (out_energy) on the fused path.  W is passed
 * pre-split: W_hi = W & 0xFFFFE000 (bitwise), W_lo = W - W_hi.  K % 32 == 0, N % 64 == 0,
 * pitches % 4 == 0, 16-byte aligned pointers.
 * ------------------------------------------------------------------------------------------- */
int hn_gemm_tf32x3(const float *A, int64_t M, int64_t K, int64_t lda, const float *W_hi, const float *W_lo,
                   int64_t N, const float *bias, float *C, int64_t ldc, void *stream);
/* Same GEMM with a fused epilogue (ScaledSiLU of rmnet.py:110-117 and its derivative):
 *   mode 0: C = A.W^T + bias
 *   mode 1: pre = A.W^T + bias;  C2 = pre (optional, row pitch ldc2);  C = silu(pre)/0.6
 *   mode 2: C = (A.W^T) * d[silu(z)/0.6]/dz at z = aux[row][col] (row pitch ld_aux) -- backward through the activation */
int hn_gemm_tf32x3_ex(const float *A, int64_t M, int64_t K, int64_t lda, const float *W_hi, const float *W_lo,
                      int64_t N, const float *bias, float *C, int64_t ldc, int32_t mode, const float *aux,
                      int64_t ld_aux, float *C2, int64_t ldc2, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused element-wise stages of the node update (HermNet/rmnet.py:24-32 residuals of PaiNNModule.forward and
 * rmnet.py:94-107 PaiNNUpdate.forward) and their hand-written backward; n rows, F channels, fp32:
 *   pre : xcat[:,0:F] = (x + dx)/sqrt2 (xcat row pitch 2F);  vecp = vec + dvec      (dx / dvec row pitches given)
 *   mid : v12 = [v1 v2] [n,3,2F]:  vdot = sum_k v1*v2/sqrt(F);  xcat[:,F:2F] = sqrt(sum_k v2^2 + 1e-8)
 *   post: a = [a1 a2 a3] [n,3F]:   x_out = xcat[:,0:F] + (a1 + a2*vdot)/sqrt2;  vec_out = vecp + a3*v1
 *   post_bwd: g_a, g_vdot, g_v12[...,0:F] from (g_x, g_vec);   mid_bwd: completes g_v12 from (g_vdot, g_cat[:,F:2F]);
 *   pre_bwd:  g_x (= g_dx) = (g_xn + g_cat[:,0:F])/sqrt2;  g_vec (= g_dvec) = g_vecn + g_vecp
 * ------------------------------------------------------------------------------------------- */
int hn_node_pre(int64_t n, int32_t F, const float *x, const float *dx, int64_t ld_dx, const float *vec,
                const float *dvec, int64_t ld_dvec, float *xcat, float *vecp, void *stream);
int hn_node_mid(int64_t n, int32_t F, const float *v12, float *vdot, float *xcat, void *stream);
int hn_node_post(int64_t n, int32_t F, const float *xcat, const float *a, const float *vdot, const float *vecp,
                 const float *v12, float *x_out, float *vec_out, void *stream);
int hn_node_post_bwd(int64_t n, int32_t F, const float *g_x, const float *g_vec, const float *a, const float *vdot,
                     const float *v12, float *g_a, float *g_vdot, float *g_v12, void *stream);
int hn_node_mid_bwd(int64_t n, int32_t F, const float *g_vdot, const float *g_cat, const float *v12, const float *vn,
                    int64_t ld_vn, float *g_v12, void *stream);
int hn_node_pre_bwd(int64_t n, int32_t F, const float *g_xn, const float *g_cat, const float *g_vecn,
                    const float *g_vecp, float *g_x, float *g_vec, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HERMNET_B200_H */
