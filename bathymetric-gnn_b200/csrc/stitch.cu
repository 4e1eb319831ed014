// Tile scatter + overlapping-tile stitch on the device (sm_100a).
//
// Replaces GraphBuilder.graph_to_grid (data/graph_construction.py:471-505) as used by
// BathymetricPipeline._process_tile (models/pipeline.py:277-314), TileManager.merge_tile
// (data/tiling.py:242-258), TileMerger.add_tile / finalize (:384-454) and the post-stitch
// fill + _apply_corrections (models/pipeline.py:196-211, 316-349).
//
// Compiled with -fmad=false: `acc += f32(v*w)` must round like NumPy.
// All kernels are elementwise over cells: HBM-bound, grid-stride, 4 cells in flight per thread.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int ST_THREADS = 256;

inline unsigned grid_for(int64_t n) {
    int64_t b = ceil_div64(n, ST_THREADS);
    const int64_t cap = 148 * 16;
    return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

__global__ void __launch_bounds__(ST_THREADS)
k_scatter_tile(const int32_t* __restrict__ node_id, int64_t n, const int64_t* __restrict__ cls,
               const float* __restrict__ conf, const float* __restrict__ corr, const float* __restrict__ lstd,
               float floor_, float* __restrict__ gcls, float* __restrict__ gconf, float* __restrict__ gcorr) {
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) {
        const int id = __ldg(node_id + t);
        const bool v = id >= 0;
        if (gcls) gcls[t] = (v && cls) ? (float)__ldg(cls + id) : 0.0f;
        if (gconf) gconf[t] = (v && conf) ? __ldg(conf + id) : 0.0f;
        if (gcorr) {
            const float nc = (v && corr) ? __ldg(corr + id) : 0.0f;
            const float ls = (v && lstd) ? __ldg(lstd + id) : 0.0f;
            gcorr[t] = nc * fmaxf(ls, floor_);           // pipeline.py:306-307
        }
    }
}

__global__ void __launch_bounds__(ST_THREADS)
k_to_grid(const int32_t* __restrict__ node_id, int64_t n, const float* __restrict__ vals, float fill,
          float* __restrict__ grid) {
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) {
        const int id = __ldg(node_id + t);
        grid[t] = id >= 0 ? __ldg(vals + id) : fill;
    }
}

template <typename T>
__global__ void __launch_bounds__(ST_THREADS) k_fill(T* __restrict__ p, int64_t n, T v) {
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) p[t] = v;
}

// accumulator planes are float32 (reference default) or float64 (TileMerger.initialize(dtypes=...), tiling.py:370-373):
// bit k of bgnn_survey::f64_mask says cont[k] is a double plane, bit 4 the discrete plane.  NumPy semantics: the fp32
// product / tile value is widened on the add / store, the weight sums stay fp32.
__device__ __forceinline__ void acc_add(const bgnn_survey& s, int k, int64_t g, float contrib) {
    if ((s.f64_mask >> k) & 1) {
        double* p = (double*)s.cont[k];
        double a = p[g];
        if (isnan(a)) a = 0.0;
        p[g] = a + (double)contrib;
    } else {
        float a = s.cont[k][g];
        if (isnan(a)) a = 0.0f;                                              // first write (:251-252)
        s.cont[k][g] = a + contrib;                                          // :258 (separate mul, add)
    }
}
__device__ __forceinline__ bool disc_isnan(const bgnn_survey& s, int64_t g) {
    return (s.f64_mask >> 4) & 1 ? isnan(((const double*)s.discrete)[g]) : isnan(s.discrete[g]);
}
__device__ __forceinline__ void disc_store(const bgnn_survey& s, int64_t g, float v) {
    if ((s.f64_mask >> 4) & 1) ((double*)s.discrete)[g] = (double)v; else s.discrete[g] = v;
}

struct AddArgs {
    bgnn_survey s;
    int64_t r0, c0;
    int h, w;
    const float* row_w;
    const float* col_w;
    const float* cont[4];
    int wsum_first[4];          // 1 if wsum[k] is not an alias of an earlier plane
    const float* disc;
    const float* conf;
};

__global__ void __launch_bounds__(ST_THREADS) k_add_tile(const AddArgs a) {
    const int64_t n = (int64_t)a.h * a.w;
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) {
        const int i = (int)(t / a.w), j = (int)(t - (int64_t)i * a.w);
        const int64_t g = (a.r0 + i) * a.s.cols + (a.c0 + j);
        const float wt = __ldg(a.row_w + i) * __ldg(a.col_w + j);          // np.outer in fp32 (tiling.py:309-311)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= a.s.n_cont || !a.cont[k]) continue;
            const float v = __ldg(a.cont[k] + t);
            if (!isfinite(v)) continue;                                      // tiling.py:247
            if (a.wsum_first[k]) a.s.wsum[k][g] += wt;                       // :255
            acc_add(a.s, k, g, v * wt);
        }
        if (a.disc && a.conf && a.s.discrete && a.s.tracker) {              // tiling.py:408-420
            const float dv = __ldg(a.disc + t), cf = __ldg(a.conf + t);
            if (isfinite(dv) && (cf > a.s.tracker[g] || disc_isnan(a.s, g))) {
                disc_store(a.s, g, dv);
                a.s.tracker[g] = cf;
            }
        }
    }
}

// Scatter fused into the accumulate: the per-node outputs of one tile go straight into the survey accumulators through
// the tile's node-id grid, without materialising the [3, h, w] tile grids (graph_to_grid x4 of pipeline.py:278-304 +
// merger.add_tile).  Channel order: cont[0] = confidence, cont[1] = correction (denormalised by max(local_std, floor),
// pipeline.py:306-307), discrete = class.  Cells that are invalid in the tile carry 0.0 like the reference's
// fill_value=0.0 grids (finite, so they take part in the blend).
struct NodeArgs {
    bgnn_survey s;
    int64_t r0, c0;
    int h, w;
    const float* row_w;
    const float* col_w;
    const int32_t* node_id;
    const int64_t* cls;
    const float* conf;
    const float* corr;
    const float* lstd;
    float floor_;
    int shared_wsum;
};

__global__ void __launch_bounds__(ST_THREADS) k_add_tile_nodes(const NodeArgs a) {
    const int64_t n = (int64_t)a.h * a.w;
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) {
        const int i = (int)(t / a.w), j = (int)(t - (int64_t)i * a.w);
        const int64_t g = (a.r0 + i) * a.s.cols + (a.c0 + j);
        const float wt = __ldg(a.row_w + i) * __ldg(a.col_w + j);
        const int id = __ldg(a.node_id + t);
        const bool v = id >= 0;
        const float cf = v ? __ldg(a.conf + id) : 0.0f;
        const float cl = v ? (float)__ldg(a.cls + id) : 0.0f;
        float cr = 0.0f;
        if (a.corr) cr = (v ? __ldg(a.corr + id) : 0.0f) * fmaxf(v ? __ldg(a.lstd + id) : 0.0f, a.floor_);
        a.s.wsum[0][g] += wt;
        acc_add(a.s, 0, g, cf * wt);
        if (a.s.n_cont > 1) {
            if (!a.shared_wsum) a.s.wsum[1][g] += wt;
            acc_add(a.s, 1, g, cr * wt);
        }
        if (a.s.discrete && a.s.tracker && (cf > a.s.tracker[g] || disc_isnan(a.s, g))) {
            disc_store(a.s, g, cl);
            a.s.tracker[g] = cf;
        }
    }
}

// TileManager.merge_tile WITHOUT a weight grid (tiling.py:259-272): where both are finite
// out = out * (1 - w) + tile * w (fp32, separate ops), where only the tile is finite out = tile.
__global__ void __launch_bounds__(ST_THREADS)
k_blend_tile(float* __restrict__ out, int64_t cols, int64_t r0, int64_t c0, int h, int w, const float* __restrict__ row_w,
             const float* __restrict__ col_w, const float* __restrict__ tile) {
    const int64_t n = (int64_t)h * w;
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) {
        const int i = (int)(t / w), j = (int)(t - (int64_t)i * w);
        const int64_t g = (r0 + i) * cols + (c0 + j);
        const float v = __ldg(tile + t);
        if (!isfinite(v)) continue;
        const float wt = __ldg(row_w + i) * __ldg(col_w + j);
        const float cur = out[g];
        out[g] = isfinite(cur) ? cur * (1.0f - wt) + v * wt : v;
    }
}

__global__ void __launch_bounds__(ST_THREADS) k_finalize(const bgnn_survey s) {
    const int64_t n = s.rows * s.cols;
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= s.n_cont) continue;
            const float w = s.wsum[k][t];
            if (!(w > 0.0f)) continue;                                       // tiling.py:289-292
            if ((s.f64_mask >> k) & 1) ((double*)s.cont[k])[t] = ((double*)s.cont[k])[t] / (double)w;
            else s.cont[k][t] = s.cont[k][t] / w;
        }
    }
}

__global__ void __launch_bounds__(ST_THREADS)
k_apply(const float* __restrict__ depth, const uint8_t* __restrict__ valid, int64_t n, float* __restrict__ cls,
        float* __restrict__ conf, float* __restrict__ corr, float* __restrict__ cleaned, float thr) {
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) {
        const bool v = valid[t] != 0;
        float c = cls[t], cf = conf[t], cr = corr[t];
        if (v && isnan(c)) {                                                 // pipeline.py:198-207
            c = 0.0f; cf = 0.0f; cr = 0.0f;
            cls[t] = c; conf[t] = cf; corr[t] = cr;
        }
        const float d = depth[t];
        if (cleaned) cleaned[t] = (c == 2.0f && cf > thr && v) ? d - cr : d; // pipeline.py:337-342
    }
}

// finalize + unprocessed-cell fill + _apply_corrections of a row range in ONE pass (the survey path emits rows as soon as
// every tile touching them has been accumulated): tiling.py:289-292, pipeline.py:196-211, 316-349.  Same fp32 operations
// as k_finalize followed by k_apply (bit-identical); valid = isfinite(depth) && depth != nodata (loaders.py:58-69) is
// derived here instead of being read from a mask plane; the class plane leaves as uint8 (255 = NaN / never written).
__global__ void __launch_bounds__(ST_THREADS)
k_survey_emit(const bgnn_survey s, const float* __restrict__ depth, int use_nodata, float nodata, float thr,
              float* __restrict__ cleaned, uint8_t* __restrict__ cls_u8, int64_t pitch) {
    const int64_t n = s.rows * s.cols;
    for (int64_t q = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; q < n; q += (int64_t)gridDim.x * ST_THREADS) {
        int64_t t = q;
        if (pitch != s.cols) { const int64_t r = q / s.cols; t = r * pitch + (q - r * s.cols); }      // a column block of wider planes
        float cf = s.cont[0][t];
        float cr = s.n_cont > 1 ? s.cont[1][t] : 0.0f;
        const float w0 = s.wsum[0][t];
        if (w0 > 0.0f) cf = cf / w0;
        if (s.n_cont > 1) {
            const float w1 = s.wsum[1][t];
            if (w1 > 0.0f) cr = cr / w1;
        }
        float c = s.discrete[t];
        const float d = depth[t];
        const bool v = isfinite(d) && !(use_nodata && d == nodata);
        if (v && isnan(c)) { c = 0.0f; cf = 0.0f; cr = 0.0f; }
        s.cont[0][t] = cf;
        if (s.n_cont > 1) s.cont[1][t] = cr;
        s.discrete[t] = c;
        if (cls_u8) cls_u8[t] = isnan(c) ? (uint8_t)255 : (uint8_t)c;
        if (cleaned) cleaned[t] = (c == 2.0f && cf > thr && v) ? d - cr : d;
    }
}

}  // namespace

extern "C" int bgnn_survey_emit(const bgnn_survey* s, const float* depth, int32_t use_nodata, float nodata,
                                float auto_correct_threshold, float* cleaned, uint8_t* class_u8, int64_t pitch, void* stream) {
    BGNN_CHECK_ARG(s && depth, "null argument");
    BGNN_CHECK_ARG(s->n_cont >= 1 && s->n_cont <= 2 && s->cont[0] && s->wsum[0] && s->discrete && s->f64_mask == 0,
                   "survey must carry float32 confidence [+ correction] and class planes");
    BGNN_CHECK_ARG(s->n_cont == 1 || (s->cont[1] && s->wsum[1]), "null correction plane");
    if (s->rows * s->cols <= 0) return BGNN_OK;
    ProfScope ps(77, (cudaStream_t)stream);
    BGNN_CHECK_ARG(pitch == 0 || pitch >= s->cols, "pitch smaller than the block width");
    k_survey_emit<<<grid_for(s->rows * s->cols), ST_THREADS, 0, (cudaStream_t)stream>>>(*s, depth, use_nodata, nodata,
                                                                                     auto_correct_threshold, cleaned, class_u8,
                                                                                     pitch > 0 ? pitch : s->cols);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_scatter_tile(const int32_t* node_id, int64_t n_cells, const int64_t* cls, const float* conf,
                                 const float* corr, const float* local_std, float norm_floor, float* class_grid,
                                 float* conf_grid, float* corr_grid, void* stream) {
    BGNN_CHECK_ARG(node_id || n_cells == 0, "node_id is null");
    if (n_cells <= 0) return BGNN_OK;
    ProfScope ps(70, (cudaStream_t)stream);
    k_scatter_tile<<<grid_for(n_cells), ST_THREADS, 0, (cudaStream_t)stream>>>(node_id, n_cells, cls, conf, corr, local_std,
                                                                               norm_floor, class_grid, conf_grid, corr_grid);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_graph_to_grid(const int32_t* node_id, int64_t n_cells, const float* values, float fill, float* grid,
                                  void* stream) {
    BGNN_CHECK_ARG((node_id && grid) || n_cells == 0, "null argument");
    if (n_cells <= 0) return BGNN_OK;
    ProfScope ps(74, (cudaStream_t)stream);
    k_to_grid<<<grid_for(n_cells), ST_THREADS, 0, (cudaStream_t)stream>>>(node_id, n_cells, values, fill, grid);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_stitch_init(const bgnn_survey* s, void* stream) {
    BGNN_CHECK_ARG(s && s->rows >= 0 && s->cols >= 0 && s->n_cont >= 0 && s->n_cont <= 4, "bad survey");
    const int64_t n = s->rows * s->cols;
    if (n == 0) return BGNN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    for (int k = 0; k < s->n_cont; ++k) {
        BGNN_CHECK_ARG(s->cont[k] && s->wsum[k], "null accumulator plane %d", k);
        if ((s->f64_mask >> k) & 1) k_fill<double><<<grid_for(n), ST_THREADS, 0, st>>>((double*)s->cont[k], n, nan(""));
        else k_fill<float><<<grid_for(n), ST_THREADS, 0, st>>>(s->cont[k], n, nanf(""));
        k_fill<float><<<grid_for(n), ST_THREADS, 0, st>>>(s->wsum[k], n, 0.0f);
    }
    if (s->discrete) {
        if ((s->f64_mask >> 4) & 1) k_fill<double><<<grid_for(n), ST_THREADS, 0, st>>>((double*)s->discrete, n, nan(""));
        else k_fill<float><<<grid_for(n), ST_THREADS, 0, st>>>(s->discrete, n, nanf(""));
    }
    if (s->tracker) k_fill<float><<<grid_for(n), ST_THREADS, 0, st>>>(s->tracker, n, -1.0f);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_stitch_add_tile(const bgnn_survey* s, int64_t r0, int64_t c0, int32_t h, int32_t w,
                                    const float* row_w, const float* col_w, const float* const* cont_tiles,
                                    const float* discrete_tile, const float* conf_tile, void* stream) {
    BGNN_CHECK_ARG(s && row_w && col_w, "null argument");
    BGNN_CHECK_ARG(h >= 0 && w >= 0 && r0 >= 0 && c0 >= 0 && r0 + h <= s->rows && c0 + w <= s->cols,
                   "tile [%lld:%lld, %lld:%lld] outside survey %lldx%lld", (long long)r0, (long long)(r0 + h),
                   (long long)c0, (long long)(c0 + w), (long long)s->rows, (long long)s->cols);
    if ((int64_t)h * w == 0) return BGNN_OK;
    AddArgs a;
    a.s = *s; a.r0 = r0; a.c0 = c0; a.h = h; a.w = w; a.row_w = row_w; a.col_w = col_w;
    for (int k = 0; k < 4; ++k) {
        a.cont[k] = (cont_tiles && k < s->n_cont) ? cont_tiles[k] : nullptr;
        a.wsum_first[k] = 1;
        for (int q = 0; q < k; ++q)
            if (k < s->n_cont && a.cont[q] && s->wsum[q] == s->wsum[k]) a.wsum_first[k] = 0;
    }
    a.disc = discrete_tile; a.conf = conf_tile;
    ProfScope ps(71, (cudaStream_t)stream);
    k_add_tile<<<grid_for((int64_t)h * w), ST_THREADS, 0, (cudaStream_t)stream>>>(a);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_stitch_add_tile_nodes(const bgnn_survey* s, int64_t r0, int64_t c0, int32_t h, int32_t w,
                                          const float* row_w, const float* col_w, const int32_t* node_id,
                                          const int64_t* cls, const float* conf, const float* corr,
                                          const float* local_std, float norm_floor, void* stream) {
    BGNN_CHECK_ARG(s && row_w && col_w && node_id && cls && conf, "null argument");
    BGNN_CHECK_ARG(s->n_cont >= 1 && s->n_cont <= 2 && s->cont[0] && s->wsum[0], "survey must carry confidence [+ correction] planes");
    BGNN_CHECK_ARG(s->n_cont == 1 || (s->cont[1] && s->wsum[1]), "null correction plane");
    BGNN_CHECK_ARG(!corr || local_std, "correction needs local_std");
    BGNN_CHECK_ARG(h >= 0 && w >= 0 && r0 >= 0 && c0 >= 0 && r0 + h <= s->rows && c0 + w <= s->cols,
                   "tile [%lld:%lld, %lld:%lld] outside survey %lldx%lld", (long long)r0, (long long)(r0 + h),
                   (long long)c0, (long long)(c0 + w), (long long)s->rows, (long long)s->cols);
    if ((int64_t)h * w == 0) return BGNN_OK;
    NodeArgs a;
    a.s = *s; a.r0 = r0; a.c0 = c0; a.h = h; a.w = w; a.row_w = row_w; a.col_w = col_w;
    a.node_id = node_id; a.cls = cls; a.conf = conf; a.corr = corr; a.lstd = local_std; a.floor_ = norm_floor;
    a.shared_wsum = s->n_cont > 1 && s->wsum[0] == s->wsum[1];
    ProfScope ps(78, (cudaStream_t)stream);
    k_add_tile_nodes<<<grid_for((int64_t)h * w), ST_THREADS, 0, (cudaStream_t)stream>>>(a);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_stitch_blend_tile(float* out, int64_t rows, int64_t cols, int64_t r0, int64_t c0, int32_t h, int32_t w,
                                      const float* row_w, const float* col_w, const float* tile, void* stream) {
    BGNN_CHECK_ARG(out && row_w && col_w && tile, "null argument");
    BGNN_CHECK_ARG(h >= 0 && w >= 0 && r0 >= 0 && c0 >= 0 && r0 + h <= rows && c0 + w <= cols, "tile outside the grid");
    if ((int64_t)h * w == 0) return BGNN_OK;
    k_blend_tile<<<grid_for((int64_t)h * w), ST_THREADS, 0, (cudaStream_t)stream>>>(out, cols, r0, c0, h, w, row_w, col_w, tile);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_stitch_finalize(const bgnn_survey* s, void* stream) {
    BGNN_CHECK_ARG(s, "null survey");
    if (s->rows * s->cols == 0 || s->n_cont == 0) return BGNN_OK;
    ProfScope ps(72, (cudaStream_t)stream);
    k_finalize<<<grid_for(s->rows * s->cols), ST_THREADS, 0, (cudaStream_t)stream>>>(*s);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_apply_corrections(const float* depth, const uint8_t* valid, int64_t n_cells, float* classification,
                                      float* confidence, float* correction, float* cleaned, float thr, void* stream) {
    BGNN_CHECK_ARG((depth && valid && classification && confidence && correction) || n_cells == 0, "null argument");
    if (n_cells <= 0) return BGNN_OK;
    ProfScope ps(73, (cudaStream_t)stream);
    k_apply<<<grid_for(n_cells), ST_THREADS, 0, (cudaStream_t)stream>>>(depth, valid, n_cells, classification, confidence,
                                                                       correction, cleaned, thr);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}
