// Native VR-BAG post-processing on the packed ragged cell space (sm_100a) -- SURVEY.md 8(f) rank 2, the step right behind
// the ragged hot path (BASELINE configs[3]).
//
// Replaces, for a whole batch of refinement grids at once,
//   * the `apply_results` closure of scripts/inference_native.py:480-503 (corrected depth / uncertainty where
//     class == NOISE && valid && confidence >= threshold, plus the run statistics), and
//   * SidecarBuilder.add_refinement_results (data/vr_bag.py:695-778): a 4-deep Python loop that paints every refinement cell
//     as a scale x scale block into the sidecar raster, south-up rows flipped to the north-up raster, LATER grids overwriting
//     earlier ones, the valid mask only ever set.
//
// Compiled with -fmad=false like stitch.cu (`depth - corr`, `unc * (2 - conf)` round like NumPy).  Elementwise over packed
// cells: HBM-bound, grid-stride.  The sidecar's "last writer wins" is made deterministic with an owner plane: pass 1 takes
// atomicMax(owner[pixel], sequence number of the grid), pass 2 lets only the owner write -- bit-identical to the
// reference's sequential loop for any overlap pattern.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int VP_THREADS = 256;

inline unsigned vp_grid(int64_t n) {
    int64_t b = ceil_div64(n, VP_THREADS);
    const int64_t cap = 148 * 16;
    return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

__global__ void __launch_bounds__(VP_THREADS)
k_vr_apply(const float* __restrict__ depth, const float* __restrict__ unc, int64_t n, float nodata,
           const float* __restrict__ cls, const float* __restrict__ conf, const float* __restrict__ corr, float thr,
           float* __restrict__ out_depth, float* __restrict__ out_unc, unsigned long long* __restrict__ counts,
           double* __restrict__ conf_sum) {
    unsigned long long n_corr = 0, n_noise = 0, n_valid = 0;
    double csum = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * VP_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * VP_THREADS) {
        const float d = __ldg(depth + t), u = unc ? __ldg(unc + t) : 0.f;
        const float c = __ldg(conf + t);
        const bool valid = (d != nodata) && isfinite(d);                       // RefinementGrid.valid_mask, vr_bag.py:90-93
        const bool noise = (__ldg(cls + t) == 2.0f) && valid;                  // inference_native.py:487
        const bool apply = noise && (c >= thr);                                // :488-489  (>=, unlike the tiled path)
        float od = d, ou = u;
        if (apply) {
            od = d - __ldg(corr + t);                                          // :492
            const float scale = 2.0f - c;                                      // :494
            ou = u * scale;                                                    // :495
        }
        out_depth[t] = od;
        if (out_unc) out_unc[t] = ou;
        n_corr += apply; n_noise += noise; n_valid += valid;
        if (valid) csum += (double)c;                                          // :502
    }
    // block reduction of the statistics
    __shared__ unsigned long long s_cnt[3][VP_THREADS / 32];
    __shared__ double s_sum[VP_THREADS / 32];
    for (int o = 16; o; o >>= 1) {
        n_corr += __shfl_xor_sync(0xffffffffu, n_corr, o);
        n_noise += __shfl_xor_sync(0xffffffffu, n_noise, o);
        n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
        csum += __shfl_xor_sync(0xffffffffu, csum, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_cnt[0][warp] = n_corr; s_cnt[1][warp] = n_noise; s_cnt[2][warp] = n_valid; s_sum[warp] = csum; }
    __syncthreads();
    if (threadIdx.x == 0 && counts) {
        unsigned long long a = 0, b = 0, c = 0;
        double s = 0.0;
        for (int w = 0; w < VP_THREADS / 32; ++w) { a += s_cnt[0][w]; b += s_cnt[1][w]; c += s_cnt[2][w]; s += s_sum[w]; }
        if (a) atomicAdd(counts + 0, a);
        if (b) atomicAdd(counts + 1, b);
        if (c) atomicAdd(counts + 2, c);
        if (conf_sum && s != 0.0) atomicAdd(conf_sum, s);
    }
}

// grid g owning packed cell t: largest g with cell_off[g] <= t
__device__ __forceinline__ int find_grid(const int64_t* __restrict__ cell_off, int n_grids, int64_t t) {
    int lo = 0, hi = n_grids - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(cell_off + mid) <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <bool WRITE>
__global__ void __launch_bounds__(VP_THREADS)
k_sidecar(bgnn_sidecar sc, int n_grids, const int64_t* __restrict__ cell_off, const int32_t* __restrict__ desc, int64_t n,
          const float* __restrict__ cls, const float* __restrict__ conf, const float* __restrict__ corr,
          const float* __restrict__ depth, float nodata) {
    for (int64_t t = (int64_t)blockIdx.x * VP_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * VP_THREADS) {
        const int g = find_grid(cell_off, n_grids, t);
        const int32_t* ds = desc + (int64_t)g * BGNN_SIDECAR_DESC;
        const int rows = __ldg(ds + 0), cols = __ldg(ds + 1), orow = __ldg(ds + 2), ocol = __ldg(ds + 3);
        const int scale = __ldg(ds + 4), seq = __ldg(ds + 5);
        const int64_t local = t - __ldg(cell_off + g);
        const int r = (int)(local / cols), c = (int)(local % cols);
        // refinement row 0 is the SOUTH edge, raster row 0 the north: vr_bag.py:758-760
        const int64_t pr0 = (int64_t)orow + (int64_t)(rows - 1 - r) * scale, pc0 = (int64_t)ocol + (int64_t)c * scale;
        float vc = 0.f, vf = 0.f, vr = 0.f;
        bool valid = false;
        if (WRITE) {
            vc = __ldg(cls + t); vf = __ldg(conf + t); vr = __ldg(corr + t);
            const float d = __ldg(depth + t);
            valid = (d != nodata) && isfinite(d);
        }
        for (int dr = 0; dr < scale; ++dr)
            for (int dc = 0; dc < scale; ++dc) {
                const int64_t pr = pr0 + dr, pc = pc0 + dc;
                if (pr < 0 || pr >= sc.rows || pc < 0 || pc >= sc.cols) continue;           // :766
                const int64_t pix = pr * sc.cols + pc;
                if (!WRITE) {
                    atomicMax(sc.owner + pix, seq);
                } else {
                    if (sc.owner[pix] == seq) {                                              // the last writer of the sequential loop
                        sc.classification[pix] = vc; sc.confidence[pix] = vf; sc.correction[pix] = vr;
                    }
                    if (valid) sc.valid_mask[pix] = 1.0f;                                    // :770-771: only ever set
                }
            }
    }
}

__global__ void __launch_bounds__(VP_THREADS) k_sidecar_init(bgnn_sidecar sc) {
    const int64_t n = sc.rows * sc.cols;
    const float nanv = nanf("");
    for (int64_t t = (int64_t)blockIdx.x * VP_THREADS + threadIdx.x; t < n; t += (int64_t)gridDim.x * VP_THREADS) {
        sc.classification[t] = nanv; sc.confidence[t] = nanv; sc.correction[t] = nanv;       // vr_bag.py:629-632
        sc.valid_mask[t] = 0.0f;
        sc.owner[t] = -1;
    }
}

}  // namespace

extern "C" int bgnn_vr_apply_results(const float* depth, const float* uncertainty, int64_t n_cells, float nodata,
                                     const float* classification, const float* confidence, const float* correction,
                                     float auto_correct_threshold, float* out_depth, float* out_uncertainty,
                                     uint64_t* counts3, double* confidence_sum, void* stream) {
    BGNN_CHECK_ARG(n_cells >= 0, "negative size");
    if (n_cells == 0) return BGNN_OK;
    BGNN_CHECK_ARG(depth && classification && confidence && correction && out_depth, "null argument");
    BGNN_CHECK_ARG((uncertainty == nullptr) == (out_uncertainty == nullptr), "uncertainty in/out must both be given or both be null");
    ProfScope ps(75, (cudaStream_t)stream);
    k_vr_apply<<<vp_grid(n_cells), VP_THREADS, 0, (cudaStream_t)stream>>>(depth, uncertainty, n_cells, nodata, classification,
                                                                         confidence, correction, auto_correct_threshold, out_depth,
                                                                         out_uncertainty, (unsigned long long*)counts3, confidence_sum);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_sidecar_init(const bgnn_sidecar* sc, void* stream) {
    BGNN_CHECK_ARG(sc && sc->rows >= 0 && sc->cols >= 0, "bad sidecar");
    if (sc->rows * sc->cols == 0) return BGNN_OK;
    BGNN_CHECK_ARG(sc->classification && sc->confidence && sc->correction && sc->valid_mask && sc->owner, "null sidecar plane");
    k_sidecar_init<<<vp_grid(sc->rows * sc->cols), VP_THREADS, 0, (cudaStream_t)stream>>>(*sc);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_sidecar_add(const bgnn_sidecar* sc, int32_t n_grids, const int64_t* cell_off, const int32_t* desc,
                                int64_t n_cells, const float* classification, const float* confidence, const float* correction,
                                const float* depth, float nodata, void* stream) {
    BGNN_CHECK_ARG(sc && n_grids >= 0 && n_cells >= 0, "bad argument");
    if (n_grids == 0 || n_cells == 0 || sc->rows * sc->cols == 0) return BGNN_OK;
    BGNN_CHECK_ARG(cell_off && desc && classification && confidence && correction && depth, "null argument");
    BGNN_CHECK_ARG(sc->classification && sc->confidence && sc->correction && sc->valid_mask && sc->owner, "null sidecar plane");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(76, st);
    k_sidecar<false><<<vp_grid(n_cells), VP_THREADS, 0, st>>>(*sc, n_grids, cell_off, desc, n_cells, classification, confidence,
                                                             correction, depth, nodata);
    k_sidecar<true><<<vp_grid(n_cells), VP_THREADS, 0, st>>>(*sc, n_grids, cell_off, desc, n_cells, classification, confidence,
                                                            correction, depth, nodata);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}
