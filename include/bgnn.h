/*
 * bgnn.h -- C ABI of libbgnn.so: the B200 (sm_100a) implementation of the
 * Bathymetric-GNN per-tile inference hot path
 *     spatial-graph build -> 4x edge-aware GAT forward + heads -> tile scatter / stitch.
 *
 * The reference (grant-froelich/Bathymetric-GNN) is pure Python and has NO FFI /
 * plugin interface for this path; the boundary it offers is the
 * Python class surface.  Each entry point below therefore names the reference
 * Python function it replaces (file:line under the reference tree); the ctypes
 * binding a maintainer adds on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer marked [dev] is device memory owned by the CALLER; the library
 *     allocates nothing on the data path (only the opaque model handle owns memory);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises except the functions documented as "host-synchronous";
 *   - return value: 0 = ok, negative = error (BGNN_E_*); text via bgnn_last_error()
 *     (thread-local);
 *   - inputs are never modified;
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails.
 */
#ifndef BGNN_H_
#define BGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGNN_VERSION 100            /* 0.1.0 */

#define BGNN_OK            0
#define BGNN_E_INVALID    -1        /* bad argument (maps to Python ValueError)   */
#define BGNN_E_CUDA       -2        /* CUDA runtime error (maps to RuntimeError)  */
#define BGNN_E_UNSUPPORTED -3       /* configuration outside the compiled kernels */
#define BGNN_E_WORKSPACE  -4        /* workspace too small                        */

#define BGNN_MAX_OFFSETS   24       /* 4-, 8-connected and the 5x5 (k=24) stencil */
#define BGNN_MAX_NODE_FEATS 8
#define BGNN_MAX_EDGE_FEATS 4

/* node feature codes: data/graph_construction.py:60-68, 288-316 */
enum { BGNN_NF_DEPTH = 0, BGNN_NF_LOCAL_MEAN, BGNN_NF_LOCAL_STD, BGNN_NF_GRAD_X, BGNN_NF_GRAD_Y,
       BGNN_NF_GRAD_MAG, BGNN_NF_CURVATURE, BGNN_NF_UNCERTAINTY };
/* edge feature codes: data/graph_construction.py:71-75, 342-369 (unknown name -> ZERO) */
enum { BGNN_EF_DISTANCE = 0, BGNN_EF_DEPTH_DIFF, BGNN_EF_SLOPE, BGNN_EF_ZERO };

int         bgnn_version(void);
const char* bgnn_last_error(void);
/* device properties needed by callers that size launches: writes {sm_count, cc_major, cc_minor} */
int         bgnn_device_info(int device, int* out3);

/* ------------------------------------------------------------------------------------------
 * Spatial-graph build.  Replaces GraphBuilder.build_graph / _build_edges /
 * _compute_node_features / _masked_local_stats / _compute_curvature / _compute_edge_features
 * (data/graph_construction.py:91-174, 176-243, 245-327, 378-432, 434-458, 329-376) and, with
 * n_grids > 1, the per-refinement-grid loop + Batch.from_data_list of
 * scripts/inference_native.py:158-179, 249-269, 312.
 *
 * A "cell space" is one or more row-major grids laid end to end:
 *   n_grids == 1 : one tile; `depth` may be a strided view into a survey (row_stride elements
 *                  between rows), so no tile copy is needed (data/tiling.py:140-178);
 *   n_grids  > 1 : ragged batch; grids are packed contiguously (row stride == cols), described
 *                  by the device arrays cell_off / rows / cols / res.
 * Nodes are the valid cells in cell-space order (== reference raster order per grid, grids in
 * batch order), node ids are the exclusive prefix count of valid cells (int32).
 * ------------------------------------------------------------------------------------------ */
typedef struct bgnn_cells {
    int32_t        n_grids;      /* >= 1 */
    int64_t        n_cells;      /* total cells (sum rows*cols)                                  */
    /* n_grids == 1 */
    int32_t        rows, cols;
    int64_t        row_stride;   /* elements between consecutive rows of depth / mask / unc      */
    double         rx, ry;       /* resolution (x, y)                                            */
    /* n_grids > 1 (all [dev]) */
    const int64_t* cell_off;     /* [n_grids+1] exclusive prefix of rows*cols                    */
    const int32_t* grid_rows;    /* [n_grids] */
    const int32_t* grid_cols;    /* [n_grids] */
    const double*  grid_res;     /* [n_grids][2] (rx, ry)                                        */
    /* data [dev] */
    const float*   depth;
    const uint8_t* mask;         /* optional; NULL => valid = isfinite(depth) [&& depth != nodata] */
    const float*   unc;          /* optional uncertainty plane                                   */
    int32_t        use_nodata;   /* when mask == NULL: also require depth != nodata             */
    float          nodata;
} bgnn_cells;

typedef struct bgnn_graph_cfg {
    int32_t n_offsets;                           /* 4, 8 or 24 */
    int32_t off_dr[BGNN_MAX_OFFSETS];            /* reference order, graph_construction.py:79-87 */
    int32_t off_dc[BGNN_MAX_OFFSETS];
    int32_t n_node_feats;                        /* columns of x (<= 8)                          */
    int32_t node_feats[BGNN_MAX_NODE_FEATS];     /* BGNN_NF_* per column                         */
    int32_t n_edge_feats;                        /* columns of edge_attr (<= 4)                  */
    int32_t edge_feats[BGNN_MAX_EDGE_FEATS];     /* BGNN_EF_* per column                         */
} bgnn_graph_cfg;

/* bytes of scratch needed by bgnn_graph_count/emit/coo for a cell space of n_cells */
size_t bgnn_graph_scratch_bytes(int64_t n_cells, int32_t n_offsets);

/* Phase 1 (host-synchronous: returns N and E): classify cells, count nodes and edges, scan.
 * totals_host[0] = N (nodes), totals_host[1] = E (directed edges, no self loops). */
int bgnn_graph_count(const bgnn_cells* cells, const bgnn_graph_cfg* cfg,
                     void* scratch, size_t scratch_bytes, int64_t* totals_host, void* stream);

/* Phase 2: emit the lean graph used by the forward pass.
 *   node_id   [dev] int32 [n_cells]   node id per cell, -1 invalid (node_index_grid, :125-126)
 *   node_cell [dev] int32 [N]         cell-space index of each node (valid_rows/cols = divmod)
 *   x         [dev] f32   [N, n_node_feats]
 *   local_std [dev] f32   [N]
 *   rowptr    [dev] int32 [N+1]       in-edge CSR; row i lists the edges (j -> i) in the
 *   csr_src   [dev] int32 [E]         reference's edge order (offset k ascending, j = i - off_k)
 *   csr_eattr [dev] f32   [E, n_edge_feats]
 *   grid_eattr [dev] f32  [n_cells, n_offsets, n_edge_feats]  (optional) the same in-edge features in a dense per-cell
 *                         layout: slot k of cell i holds the features of the edge (i - off_k -> i), zeros when that edge
 *                         does not exist; rows of invalid cells are not written.  Consumed by the tiled aggregation
 *                         (bgnn_grid_topo.grid_eattr), which can then fetch them without the rowptr indirection.
 * Any of x/local_std/csr_* / grid_eattr may be NULL to skip that output. */
int bgnn_graph_emit(const bgnn_cells* cells, const bgnn_graph_cfg* cfg,
                    void* scratch, size_t scratch_bytes,
                    int32_t* node_id, int32_t* node_cell, float* x, float* local_std,
                    int32_t* rowptr, int32_t* csr_src, float* csr_eattr, float* grid_eattr, void* stream);

/* The same second phase for a RAGGED batch whose grids are cut into band patches (bgnn_grid_topo::patches; the reference's
 * 8-connected offsets): one CTA per band with depth + validity staged in shared memory, no per-cell grid search, four
 * fp64 atans per cell instead of eight.  Writes node_id / node_cell / rowptr, x, local_std and the dense grid_eattr (the
 * CSR payload stays available through bgnn_graph_emit).  Bit-identical to bgnn_graph_emit. */
int bgnn_graph_emit_bands(const bgnn_cells* cells, const bgnn_graph_cfg* cfg, void* scratch, size_t scratch_bytes,
                          const void* patches, int32_t n_patches, int32_t* node_id, int32_t* node_cell, float* x,
                          float* local_std, int32_t* rowptr, float* grid_eattr, void* stream);

/* Optional phase 3: the reference's API-visible COO (offset-major, node-minor; :196-241) and
 * per-edge features in that order.  edge_index int64 [2,E] (row 0 = src/centre, row 1 = tgt),
 * edge_attr f32 [E, n_edge_feats].  Needs the scratch of the SAME count call and node_id. */
int bgnn_graph_coo(const bgnn_cells* cells, const bgnn_graph_cfg* cfg,
                   void* scratch, size_t scratch_bytes, const int32_t* node_id,
                   int64_t n_edges, int64_t* edge_index, float* edge_attr, void* stream);

/* ------------------------------------------------------------------------------------------
 * GAT forward.  Replaces BathymetricGNN.forward / predict (models/gnn.py:360-451):
 * LocalFeatureExtractor (:55-68) -> L x [GATConv(edge_dim) -> BatchNorm(eval) -> ReLU]
 * (:107-154, 173-186; PyG GATConv semantics) -> three heads (:191-260).
 * ------------------------------------------------------------------------------------------ */
typedef struct bgnn_model_desc {
    int32_t in_channels;        /* F (7 or 8)                      */
    int32_t hidden;             /* C, per-head width (64 / 128)    */
    int32_t layers;             /* L (4)                           */
    int32_t heads;              /* H (4); last layer uses 1 head   */
    int32_t num_classes;        /* 3                               */
    int32_t predict_correction; /* 0/1                             */
    int32_t edge_dim;           /* 0 (None) .. 4                   */
    float   bn_eps;             /* 1e-5                            */
    float   negative_slope;     /* 0.2                             */
} bgnn_model_desc;

typedef struct bgnn_model bgnn_model;   /* opaque: folded weights resident on one device */

/* number of floats bgnn_model_create expects, in this order (all row-major, fp32):
 *   feature_extractor.mlp.0.{weight[C,F],bias[C]}, feature_extractor.mlp.3.{weight[C,C],bias[C]},
 *   per layer l: lin.weight[HC_l,in_l], att_src[HC_l], att_dst[HC_l],
 *                (edge_dim>0: lin_edge.weight[HC_l,edge_dim], att_edge[HC_l]),
 *                bias[out_l], norm.weight[out_l], norm.bias[out_l], running_mean[out_l], running_var[out_l]
 *   classification_head.mlp.0.{weight[C/2,C],bias}, .mlp.3.{weight[num_classes,C/2],bias},
 *   confidence_head.mlp.0.{..}, .mlp.3.{weight[1,C/2],bias[1]},
 *   (predict_correction: correction_head.mlp.0.{..}, .mlp.3.{..})
 * with H_l = heads (l < L-1) or 1 (last), HC_l = H_l*C, in_0 = C, in_l = heads*C, out_l = HC_l
 * (concat) or C (last layer, mean over its single head). */
int64_t bgnn_model_param_floats(const bgnn_model_desc* desc);
/* host-synchronous; params_host is HOST memory */
int  bgnn_model_create(const bgnn_model_desc* desc, const float* params_host, int64_t n_floats,
                       int device, bgnn_model** out);
void bgnn_model_destroy(bgnn_model* m);

size_t bgnn_gat_workspace_bytes(const bgnn_model* m, int64_t n_nodes);

/* Compute precision of bgnn_gat_forward for this model (BASELINE configs[4] "bf16 vs fp32 tolerance"; the reference has no
 * reduced-precision path: `grep -i "autocast|bfloat|half("` over its tree is empty).
 *   BGNN_PREC_F32  (default)  fp32 parity path (tolerance: atol 2e-5 + rtol 1e-5 on logits / confidence / correction).
 *   BGNN_PREC_BF16            the per-node streams live in HBM as bf16 -- feature-extractor output, projection weights,
 *                             h = x W^T, every layer output but the last -- while every sum (GEMM accumulators in TMEM /
 *                             registers, attention logits, softmax, weighted aggregation), BatchNorm and the heads stay
 *                             fp32.  Halves the bytes the projection and aggregation kernels move.  Bounded separately:
 *                             see tests/test_gpu_bf16.py (vs the fp32 oracle and vs the oracle restating the same
 *                             rounding points).  Runs the CSR aggregation for every stencil (4-/8-/24-offset, ragged).
 * set_precision is host-synchronous the first time bf16 is requested (weights are rounded once); not thread-safe with a
 * concurrent forward on the same model. */
#define BGNN_PREC_F32  0
#define BGNN_PREC_BF16 1
int bgnn_model_set_precision(bgnn_model* m, int precision);
int bgnn_model_precision(const bgnn_model* m);

/* Optional grid topology of the node set (single tile, n_grids == 1): when given, and the stencil
 * has radius 1 (4-/8-connected), the aggregation runs as a shared-memory tiled kernel: one CTA per
 * 16x16 patch of cells stages every source row of its 18x18 halo ONCE (cp.async) instead of
 * gathering 9 rows per destination through L1.  The in-edge CSR is still required (edge features).
 * rows == 0 (node_id NULL) marks a node set WITHOUT a single grid -- a ragged batch of grids packed end to end -- and
 * only n_offsets is read: it bounds the in-degree (<= 8), which lets the CSR aggregation stage each block's source-id
 * window in shared memory (raster numbering keeps sources within a grid width of their destination). */
typedef struct bgnn_grid_topo {
    int32_t        rows, cols;
    const int32_t* node_id;                      /* [dev] int32 [rows*cols] from bgnn_graph_emit */
    int32_t        n_offsets;
    int32_t        off_dr[BGNN_MAX_OFFSETS];
    int32_t        off_dc[BGNN_MAX_OFFSETS];
    const float*   grid_eattr;                   /* [dev] optional dense edge features from bgnn_graph_emit */
    /* ragged batches (rows == 0) only: band patches planned by the host -- one entry per (grid, band of rows):
     * { int64 cell_off; int32 grid_rows, grid_cols, first_row, n_rows } with (n_rows + 2) * (grid_cols + 2) <= 324,
     * n_rows * grid_cols <= 256, n_rows <= 62.  With node_id (packed cell space) and grid_eattr set, the aggregation of a
     * ragged batch runs as the tiled kernel per band instead of the windowed CSR kernel and the CSR payload is never read. */
    const void*    patches;                      /* [dev] or NULL */
    int32_t        n_patches;
} bgnn_grid_topo;

/* 1 if bgnn_gat_forward would run EVERY layer of `m` through the tiled grid kernel for this topology (then
 * csr_src / csr_eattr are never read and may be NULL), 0 otherwise. */
int bgnn_gat_grid_path(const bgnn_model* m, const bgnn_grid_topo* topo, int64_t n_edges);

/* outputs (any may be NULL): logits/probs f32 [N,num_classes], cls int64 [N], conf f32 [N],
 * corr f32 [N], action int64 [N] (0 keep, 1 auto-correct, 2 review; gnn.py:436-445).
 * `hidden_out` (optional, f32 [N, C]) receives the backbone output x_L (testing aid). */
int bgnn_gat_forward(const bgnn_model* m, int64_t n_nodes, int64_t n_edges,
                     const float* x, const int32_t* rowptr, const int32_t* csr_src,
                     const float* csr_eattr, void* workspace, size_t workspace_bytes,
                     float* logits, float* probs, int64_t* cls, float* conf, float* corr,
                     int64_t* action, float auto_correct_threshold, float review_threshold,
                     float* hidden_out, const bgnn_grid_topo* topo /* may be NULL */, void* stream);

/* ------------------------------------------------------------------------------------------
 * GCN / GraphSAGE / GIN backbones (models/gnn.py:120-143, `gnn_type != "GAT"`).  Layers without
 * attention on the same in-edge CSR; the host layer composes a layer from these three entry points (PyG semantics:
 * oracle/pyg_restated.py).  All pointers are DEVICE pointers; rows may be strided (ld* in floats).
 * ------------------------------------------------------------------------------------------ */
enum { BGNN_AGG_GCN = 0, BGNN_AGG_MEAN = 1, BGNN_AGG_GIN = 2 };
/* out[i, 0:C] = epilogue(reduce over the in-edges of i, sources in CSR order = the reference's edge order, self term last):
 *   BGNN_AGG_GCN   sum_j x_j * (t[d_j + 1] * t[d_i + 1]) + x_i * t[d_i + 1]^2,  t = deg_inv_sqrt table (t[k] = k^-1/2 as float32,
 *                  computed by the caller exactly as the reference does), d = in-degree (PyG gcn_norm with self loops of weight 1)
 *   BGNN_AGG_MEAN  sum_j x_j / max(d_i, 1)                                       (SAGEConv, aggr = 'mean')
 *   BGNN_AGG_GIN   sum_j x_j + (1 + eps) * x_i                                   (GINConv)
 * epilogue: (+ pre_add[c]) (* mul[c] + add[c]) (ReLU), each optional (NULL / 0).  copy_self = 1 additionally writes x_i to
 * out[i, C:2C] (SAGEConv's `lin_l(mean) + lin_r(x)` as one GEMM over [mean | x]). */
int bgnn_csr_aggregate(int mode, int64_t n_nodes, int channels, const float* x, int64_t ldx, const int32_t* rowptr,
                       const int32_t* csr_src, const float* deg_inv_sqrt, int deg_table_len, float eps,
                       const float* pre_add, const float* mul, const float* add, int relu, float* out, int64_t ldo,
                       int copy_self, void* stream);
/* out[n, 0:out_features] = (ReLU)(x[n, 0:in_features] . weight[out_features, in_features]^T + bias)   (torch.nn.Linear layout) */
int bgnn_linear(const float* x, int64_t ldx, int64_t n_rows, int in_features, const float* weight, const float* bias,
                int relu, int out_features, float* out, int64_t ldo, void* stream);
/* second layers of the heads + softmax / argmax (first max) / sigmoid / action rule (models/gnn.py:191-260, 381-451) from
 * hidden = ReLU(first layers) [N, nblk * h2] (blocks: classification | confidence | correction).  Outputs may be NULL. */
int bgnn_heads_finish(const float* hidden, int64_t n_nodes, int h2, int nblk, int num_classes, const float* Wcls,
                      const float* bcls, const float* Wconf, const float* bconf, const float* Wcorr, const float* bcorr,
                      float* logits, float* probs, int64_t* cls, float* conf, float* corr, int64_t* action,
                      float auto_correct_threshold, float review_threshold, void* stream);

/* ------------------------------------------------------------------------------------------
 * Scatter + stitch.  Replaces GraphBuilder.graph_to_grid (graph_construction.py:471-505) as
 * used by BathymetricPipeline._process_tile (models/pipeline.py:277-314), TileManager.merge_tile
 * weighted branch (data/tiling.py:242-258), TileMerger.add_tile discrete branch (:408-420) and
 * finalize (:274-294, 430-454).
 * ------------------------------------------------------------------------------------------ */
/* node values -> dense tile grids (fill 0.0): class as float, confidence,
 * correction = corr * max(local_std, norm_floor).  node_id is the int32 [rows*cols] grid from
 * bgnn_graph_emit.  Any output may be NULL. */
int bgnn_scatter_tile(const int32_t* node_id, int64_t n_cells,
                      const int64_t* cls, const float* conf, const float* corr,
                      const float* local_std, float norm_floor,
                      float* class_grid, float* conf_grid, float* corr_grid, void* stream);

/* generic graph_to_grid: grid[cell] = node_id[cell] >= 0 ? values[node_id[cell]] : fill */
int bgnn_graph_to_grid(const int32_t* node_id, int64_t n_cells, const float* values, float fill,
                       float* grid, void* stream);

typedef struct bgnn_survey {
    int64_t rows, cols;
    /* all [dev], row-major [rows, cols]; NaN / 0 / -1 initialised by bgnn_stitch_init */
    float* cont[4];      /* weighted-sum accumulators of up to 4 continuous channels */
    float* wsum[4];      /* their weight sums (may alias each other when all tiles are finite) */
    int32_t n_cont;
    float* discrete;     /* classification (float, NaN = never written) */
    float* tracker;      /* confidence tracker, init -1 (tiling.py:382) */
    int32_t f64_mask;    /* bit k: cont[k] is a DOUBLE plane, bit 4: discrete is (TileMerger.initialize(dtypes=...),
                          * tiling.py:356-373); the weight sums and the tracker are always float32 like the reference's */
} bgnn_survey;

int bgnn_stitch_init(const bgnn_survey* s, void* stream);
/* add one tile: region [r0:r0+h, c0:c0+w].  row_w [h], col_w [w] are the 1-D blend tables
 * (tiling.py:313-330, computed on the host in fp64 exactly as the reference does); the 2-D
 * weight is their fp32 product.  cont_tiles[k] is the tile grid of channel k ([h,w] f32, may
 * contain NaN = skipped), discrete_tile / conf_tile drive the arg-max-confidence rule.
 * Tiles must be added in the reference's iteration order on ONE stream (fp32 sum order). */
int bgnn_stitch_add_tile(const bgnn_survey* s, int64_t r0, int64_t c0, int32_t h, int32_t w,
                         const float* row_w, const float* col_w,
                         const float* const* cont_tiles, const float* discrete_tile,
                         const float* conf_tile, void* stream);
/* The same accumulate with the per-tile scatter fused in (replaces the 4x graph_to_grid of models/pipeline.py:278-304 +
 * TileMerger.add_tile, data/tiling.py:384-428): the tile's per-node outputs are read through its node-id grid
 * (node_id [h*w], -1 = invalid -> 0.0 like fill_value=0.0) and accumulated into cont[0] = confidence,
 * cont[1] = correction * max(local_std, norm_floor) (n_cont == 1: confidence only; corr may be NULL -> 0),
 * discrete = class by arg-max confidence.  Bit-identical to bgnn_scatter_tile + bgnn_stitch_add_tile. */
int bgnn_stitch_add_tile_nodes(const bgnn_survey* s, int64_t r0, int64_t c0, int32_t h, int32_t w,
                               const float* row_w, const float* col_w, const int32_t* node_id,
                               const int64_t* cls, const float* conf, const float* corr,
                               const float* local_std, float norm_floor, void* stream);
/* TileManager.merge_tile WITHOUT a weight grid (tiling.py:259-272): out [rows, cols] f32 is blended in place,
 * out = out * (1 - w) + tile * w where both are finite, out = tile where only the tile is. */
int bgnn_stitch_blend_tile(float* out, int64_t rows, int64_t cols, int64_t r0, int64_t c0, int32_t h, int32_t w,
                           const float* row_w, const float* col_w, const float* tile, void* stream);
/* out = acc / wsum where wsum > 0 (tiling.py:289-292) */
int bgnn_stitch_finalize(const bgnn_survey* s, void* stream);

/* post-stitch: unprocessed valid cells -> class 0 / conf 0 / corr 0 (pipeline.py:196-207) and
 * cleaned = depth - corr where class == 2 && conf > thr && valid (pipeline.py:316-349).
 * cont[] order is the caller's; pass the planes explicitly. */
int bgnn_apply_corrections(const float* depth, const uint8_t* valid, int64_t n_cells,
                           float* classification, float* confidence, float* correction,
                           float* cleaned, float auto_correct_threshold, void* stream);

/* Survey path: finalize (tiling.py:289-292) + the fill of unprocessed valid cells (pipeline.py:196-207) +
 * _apply_corrections (pipeline.py:316-349) over the rows of `s` in ONE pass, in place: cont[0] = confidence and
 * cont[1] = correction leave finalised, `cleaned` ([rows, cols], may alias `depth`) receives the corrected depth and
 * `class_u8` (may be NULL) the class plane as uint8 (255 = NaN: never written and not valid).  valid = isfinite(depth)
 * && depth != nodata (data/loaders.py:58-69).  Bit-identical to bgnn_stitch_finalize + bgnn_apply_corrections.  `s`
 * describes a ROW RANGE of the survey (plane pointers offset to its first row), so rows can be emitted as soon as every
 * tile that touches them has been accumulated. */
int bgnn_survey_emit(const bgnn_survey* s, const float* depth, int32_t use_nodata, float nodata,
                     float auto_correct_threshold, float* cleaned, uint8_t* class_u8,
                     int64_t pitch /* row pitch in elements of every plane; 0 = s->cols (a COLUMN BLOCK of wider planes otherwise) */,
                     void* stream);
/* cudaMemcpy2DAsync between pinned host memory and the device (kind 1 = host->device, 2 = device->host): column blocks of
 * the survey planes go up / come down tile by tile (survey.py) */
int bgnn_memcpy2d_async(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes, size_t width_bytes,
                        size_t height, int kind, void* stream);

/* ------------------------------------------------------------------------------------------
 * Native VR-BAG post-processing on the packed ragged cell space (the step right behind the ragged
 * path).  All arrays are DEVICE pointers over the packed cells of a batch of refinement grids (grid g occupies
 * [cell_off[g], cell_off[g+1]) in row-major order, as built by bgnn_graph_count/emit with n_grids > 1).
 * ------------------------------------------------------------------------------------------ */
/* Replaces the `apply_results` closure of scripts/inference_native.py:480-503 for a whole batch:
 *   valid = depth != nodata && isfinite(depth); noise = classification == 2 && valid; apply = noise && confidence >= thr (>=);
 *   out_depth = apply ? depth - correction : depth;  out_uncertainty = apply ? uncertainty * (2 - confidence) : uncertainty.
 * counts3 (optional, uint64 [3], ACCUMULATED: zero it first): cells corrected, cells classified noise, valid cells;
 * confidence_sum (optional, f64, accumulated): sum of confidence over valid cells (stats of :499-502).
 * uncertainty / out_uncertainty may both be NULL. */
int bgnn_vr_apply_results(const float* depth, const float* uncertainty, int64_t n_cells, float nodata,
                          const float* classification, const float* confidence, const float* correction,
                          float auto_correct_threshold, float* out_depth, float* out_uncertainty,
                          uint64_t* counts3, double* confidence_sum, void* stream);

/* Sidecar raster of SidecarBuilder (data/vr_bag.py:608-632): three result planes (NaN-initialised), the valid mask and an
 * owner plane (int32, sequence number of the grid that wrote the pixel last) that makes "later grids overwrite earlier
 * ones" deterministic on the device. */
typedef struct bgnn_sidecar {
    int64_t  rows, cols;
    float*   classification;    /* [dev] f32 [rows*cols] */
    float*   confidence;
    float*   correction;
    float*   valid_mask;        /* 0 / 1 */
    int32_t* owner;             /* [dev] int32 [rows*cols] */
} bgnn_sidecar;
int bgnn_sidecar_init(const bgnn_sidecar* sc, void* stream);
/* Replaces SidecarBuilder.add_refinement_results (data/vr_bag.py:695-778) for a batch of grids: every refinement cell (r, c)
 * of grid g is painted as a scale x scale block at raster row out_row_start + (rows-1-r)*scale (refinement rows run south to
 * north), column out_col_start + c*scale, clipped to the raster; valid_mask is set where the cell is valid.
 * desc: [dev] int32 [n_grids][BGNN_SIDECAR_DESC] = {rows, cols, out_row_start, out_col_start, scale, sequence number};
 * the placement integers are computed by the host exactly as the reference does (float64, int(round(.)), :716-748);
 * sequence numbers must increase in the reference's processing order across ALL calls on one raster. */
#define BGNN_SIDECAR_DESC 6
int bgnn_sidecar_add(const bgnn_sidecar* sc, int32_t n_grids, const int64_t* cell_off, const int32_t* desc, int64_t n_cells,
                     const float* classification, const float* confidence, const float* correction,
                     const float* depth, float nodata, void* stream);

/* Kernel-selection switches of ONE model handle (the library keeps no mutable process-wide state: handles on different
 * devices / host threads are independent).  All on by default; results stay within the stated tolerances either way
 * (presplit: bit-identical).  Used by the tests to cross-check the kernel variants against each other.
 *   BGNN_OPT_TENSOR_CORES  every `x W^T` whose shape is compiled for the tcgen05 path runs as an fp16 hi/lo-split GEMM with
 *                          fp32 accumulation in TMEM (fp32-grade accuracy); off -> the fp32 FMA GEMM
 *   BGNN_OPT_PRESPLIT      a tiled aggregation hands its output to the next tcgen05 projection as fp16 hi / lo planes
 *   BGNN_OPT_WINDOW        ragged batches: windowed shared-memory aggregation (off -> warp-per-destination CSR kernel)
 *   BGNN_OPT_TILE5         5x5 stencil / bf16 mode on a single grid: radius-2 tiled aggregation (off -> CSR kernel)
 *   BGNN_OPT_TILE5_WINDOW  full 5x5 stencil: sliding register window variant of the radius-2 kernel
 *   BGNN_OPT_RAGGED_PATCHES ragged batches: tiled aggregation per band of grid rows (bgnn_grid_topo::patches; off -> windowed kernel) */
enum { BGNN_OPT_TENSOR_CORES = 0, BGNN_OPT_PRESPLIT = 1, BGNN_OPT_WINDOW = 2, BGNN_OPT_TILE5 = 3, BGNN_OPT_TILE5_WINDOW = 4,
       BGNN_OPT_RAGGED_PATCHES = 5 };
int bgnn_model_set_option(bgnn_model* m, int option, int value);
/* testing aid: C_dev[M,N] = A_dev[M,K] . W_host[N,K]^T through the tcgen05 path (host-synchronous);
 * alpha_dev (optional) [M, 2*heads] receives the per-head dot products with att_src / att_dst [N]. */
int bgnn_test_proj_umma(const float* A_dev, int64_t M, int K, const float* W_host, int N, float* C_dev,
                        const float* att_src_dev, const float* att_dst_dev, float* alpha_dev, int heads,
                        void* stream);
/* same, with A handed over as fp16 hi / lo planes (the PM_SPLIT variant; N in {64, 256}) */
int bgnn_test_proj_umma_presplit(const float* A_dev, int64_t M, int K, const float* W_host, int N, float* C_dev,
                                 const float* att_src_dev, const float* att_dst_dev, float* alpha_dev, int heads,
                                 void* stream);
/* testing aid, bf16 mode: C_dev (bf16 bit patterns, HEAD-MAJOR [heads][M][N/heads]) = A_dev[M,K] (bf16) . W_dev[N,K]^T (bf16)
 * with fp32 accumulation on the tcgen05 path; alpha_dev [M, 2*heads] fp32 from the fp32 accumulators.  Returns
 * BGNN_E_UNSUPPORTED for shapes the bf16 tensor-core kernel is not compiled for. */
int bgnn_test_proj_bf16(const uint16_t* A_dev, int64_t M, int K, const uint16_t* W_dev, int N, uint16_t* C_dev,
                        const float* att_src_dev, const float* att_dst_dev, float* alpha_dev, int heads, void* stream);

/* ------------------------------------------------------------------------------------------
 * Measurement aid (bench.py): per-kernel device timing with CUDA events recorded on the
 * launching stream around every kernel the library launches.  Off by default; the record list is mutex-guarded.
 * Slot ids: 0 gb_count, 1 gb_scan, 2 gb_ids, 3 gb_emit, 4 gb_coo, 10 mlp0, 11 mlp1,
 * 20+l projection of layer l, 40+l aggregation of layer l, 60 heads gemm, 61 heads finish,
 * 70 scatter_tile, 71 stitch_add_tile, 72 stitch_finalize, 73 apply_corrections, 74 graph_to_grid, 77 survey_emit, 78 stitch_add_tile_nodes, 80 csr_aggregate, 81 linear,
 * 75 vr_apply_results, 76 sidecar_add.
 * ------------------------------------------------------------------------------------------ */
int bgnn_profile_enable(int on);                               /* also clears recorded events */
/* waits for the recorded events; fills up to `cap` (slot, milliseconds) pairs in launch order;
 * returns the number of records (may exceed cap) or a negative error */
int bgnn_profile_collect(int cap, int32_t* slots, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* BGNN_H_ */
