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

__global__ void __launch_bounds__(ST_THREADS) k_fill(float* __restrict__ p, int64_t n, float v) {
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) p[t] = v;
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
            float acc = a.s.cont[k][g];
            if (isnan(acc)) acc = 0.0f;                                      // first write (:251-252)
            if (a.wsum_first[k]) a.s.wsum[k][g] += wt;                       // :255
            a.s.cont[k][g] = acc + v * wt;                                   // :258 (separate mul, add)
        }
        if (a.disc && a.conf && a.s.discrete && a.s.tracker) {              // tiling.py:408-420
            const float dv = __ldg(a.disc + t), cf = __ldg(a.conf + t);
            const float cur = a.s.discrete[g];
            if (isfinite(dv) && (cf > a.s.tracker[g] || isnan(cur))) {
                a.s.discrete[g] = dv;
                a.s.tracker[g] = cf;
            }
        }
    }
}

__global__ void __launch_bounds__(ST_THREADS) k_finalize(const bgnn_survey s) {
    const int64_t n = s.rows * s.cols;
    for (int64_t t = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * ST_THREADS) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= s.n_cont) continue;
            const float w = s.wsum[k][t];
            if (w > 0.0f) s.cont[k][t] = s.cont[k][t] / w;                   // tiling.py:289-292
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

}  // namespace

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
        k_fill<<<grid_for(n), ST_THREADS, 0, st>>>(s->cont[k], n, nanf(""));
        k_fill<<<grid_for(n), ST_THREADS, 0, st>>>(s->wsum[k], n, 0.0f);
    }
    if (s->discrete) k_fill<<<grid_for(n), ST_THREADS, 0, st>>>(s->discrete, n, nanf(""));
    if (s->tracker) k_fill<<<grid_for(n), ST_THREADS, 0, st>>>(s->tracker, n, -1.0f);
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
