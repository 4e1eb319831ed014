// Spatial-graph build on the device (sm_100a).
//
// Replaces data/graph_construction.py:91-458 of the reference (GraphBuilder.build_graph and
// its helpers) and, for ragged batches, the per-grid loop + Batch.from_data_list of
// scripts/inference_native.py:158-179, 249-269, 312.   See include/bgnn.h for the contract.
//
// This translation unit is compiled with -fmad=false: the feature arithmetic has to round
// exactly like NumPy (separate multiply / add), and the kernels are
// bound by their output bytes (177 B/cell), not by FLOPs.
//
// Pipeline (all thread-per-cell, 256 cells per CTA, cells in cell-space order):
//   gb_count : classify cell, neighbour bitmask -> per-CTA counts {valid, degree, per-offset}
//   gb_scan  : exclusive scan of the per-CTA table (one CTA)
//   gb_ids   : node_id grid, node_cell, rowptr          (needs the scan)
//   gb_emit  : x, local_std, in-edge CSR (src ids + edge features)   (needs the node_id grid)
//   gb_coo   : optional reference-ordered COO (offset-major) for the API-visible edge_index
#include "common.cuh"

namespace {

constexpr int GB_THREADS = 256;
constexpr int GB_WARPS = GB_THREADS / 32;
constexpr int GB_MAXCOL = BGNN_MAX_OFFSETS + 2;

struct GraphParams {
    bgnn_cells cells;
    int K;
    int dr[BGNN_MAX_OFFSETS], dc[BGNN_MAX_OFFSETS], neg[BGNN_MAX_OFFSETS];
    int F, nf[BGNN_MAX_NODE_FEATS];
    int ED, ef[BGNN_MAX_EDGE_FEATS];
    int ncol;           // K + 2
    int64_t nblocks;    // CTAs of GB_THREADS cells
};

struct Ctx {
    int H, W, r, c;
    int64_t dbase, dstride;   // depth/mask/unc index = dbase + r*dstride + c
    int64_t cbase;            // packed cell index   = cbase + r*W + c
    double rx, ry;
};

__device__ __forceinline__ void locate(const bgnn_cells& cs, int64_t t, Ctx& x) {
    if (cs.n_grids == 1) {
        x.H = cs.rows; x.W = cs.cols;
        x.r = (int)(t / cs.cols); x.c = (int)(t - (int64_t)x.r * cs.cols);
        x.dbase = 0; x.dstride = cs.row_stride; x.cbase = 0;
        x.rx = cs.rx; x.ry = cs.ry;
    } else {
        int lo = 0, hi = cs.n_grids;           // largest g with cell_off[g] <= t
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (__ldg(cs.cell_off + mid) <= t) lo = mid; else hi = mid;
        }
        int64_t off = __ldg(cs.cell_off + lo);
        x.H = __ldg(cs.grid_rows + lo); x.W = __ldg(cs.grid_cols + lo);
        int64_t l = t - off;
        x.r = (int)(l / x.W); x.c = (int)(l - (int64_t)x.r * x.W);
        x.dbase = off; x.dstride = x.W; x.cbase = off;
        x.rx = __ldg(cs.grid_res + 2 * lo); x.ry = __ldg(cs.grid_res + 2 * lo + 1);
    }
}

__device__ __forceinline__ bool valid_at(const bgnn_cells& cs, int64_t di) {
    if (cs.mask) return __ldg(cs.mask + di) != 0;
    float d = __ldg(cs.depth + di);
    return isfinite(d) && (!cs.use_nodata || d != cs.nodata);
}

__device__ __forceinline__ bool valid_rc(const bgnn_cells& cs, const Ctx& x, int r, int c) {
    if (r < 0 || r >= x.H || c < 0 || c >= x.W) return false;
    return valid_at(cs, x.dbase + (int64_t)r * x.dstride + c);
}

__device__ __forceinline__ uint32_t neighbour_mask(const GraphParams& P, const Ctx& x) {
    uint32_t nb = 0;
    for (int k = 0; k < P.K; ++k)
        if (valid_rc(P.cells, x, x.r + P.dr[k], x.c + P.dc[k])) nb |= 1u << k;
    return nb;
}

// ---------------------------------------------------------------- count / scan / ids
__global__ void __launch_bounds__(GB_THREADS) gb_count(const GraphParams P, int32_t* __restrict__ table) {
    __shared__ int s_cnt[GB_WARPS][GB_MAXCOL];
    const int64_t t = (int64_t)blockIdx.x * GB_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool v = false; uint32_t nb = 0;
    if (t < P.cells.n_cells) {
        Ctx x; locate(P.cells, t, x);
        v = valid_rc(P.cells, x, x.r, x.c);
        if (v) nb = neighbour_mask(P, x);
    }
    const unsigned bv = __ballot_sync(0xffffffffu, v);
    int deg = __popc(nb);
    for (int o = 16; o; o >>= 1) deg += __shfl_xor_sync(0xffffffffu, deg, o);
    if (lane == 0) { s_cnt[warp][0] = __popc(bv); s_cnt[warp][1] = deg; }
    for (int k = 0; k < P.K; ++k) {
        const unsigned bk = __ballot_sync(0xffffffffu, (nb >> k) & 1u);
        if (lane == 0) s_cnt[warp][2 + k] = __popc(bk);
    }
    __syncthreads();
    if (threadIdx.x < P.ncol) {
        int sum = 0;
        for (int w = 0; w < GB_WARPS; ++w) sum += s_cnt[w][threadIdx.x];
        table[(int64_t)threadIdx.x * (P.nblocks + 1) + blockIdx.x] = sum;
    }
}

// one CTA per column: exclusive scan of column blockIdx.x of table[ncol][nblocks+1]; totals land at [nblocks]
// The node / edge totals (columns 0 and 1) are ALSO summed in 64 bits into totals64[col]: the int32 table wraps once a
// cell space holds 2^31 edges (268 M valid cells at K = 8, 89 M at K = 24) and bgnn_graph_count must refuse such a space
// instead of handing a wrapped E to the caller (silent out-of-bounds CSR writes).
__global__ void __launch_bounds__(1024) gb_scan(int32_t* __restrict__ table, int64_t nblocks, int ncol,
                                                long long* __restrict__ totals64) {
    __shared__ int s_part[1024];
    __shared__ long long s_wide[32];
    const int tid = threadIdx.x;
    const int64_t per = (nblocks + 1023) / 1024;
    const int64_t b0 = min((int64_t)tid * per, nblocks), b1 = min(b0 + per, nblocks);
    for (int col = blockIdx.x; col < ncol; col += gridDim.x) {
        int32_t* colp = table + (int64_t)col * (nblocks + 1);
        int sum = 0;
        for (int64_t b = b0; b < b1; ++b) sum += colp[b];
        s_part[tid] = sum;
        if (col < 2) {                              // 64-bit total (per-thread partials stay far below 2^31)
            long long w = sum;
            for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
            if ((tid & 31) == 0) s_wide[tid >> 5] = w;
        }
        __syncthreads();
        if (col < 2 && tid < 32) {
            long long w = s_wide[tid];
            for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
            if (tid == 0) totals64[col] = w;
        }
        for (int o = 1; o < 1024; o <<= 1) {       // Hillis-Steele inclusive scan
            int v = tid >= o ? s_part[tid - o] : 0;
            __syncthreads();
            s_part[tid] += v;
            __syncthreads();
        }
        int run = s_part[tid] - sum;
        for (int64_t b = b0; b < b1; ++b) { int v = colp[b]; colp[b] = run; run += v; }
        if (tid == 1023) colp[nblocks] = s_part[1023];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(GB_THREADS) gb_ids(const GraphParams P, const int32_t* __restrict__ table,
                                                     int32_t* __restrict__ node_id, int32_t* __restrict__ node_cell,
                                                     int32_t* __restrict__ rowptr) {
    __shared__ int s_v[GB_WARPS], s_d[GB_WARPS];
    const int64_t t = (int64_t)blockIdx.x * GB_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool v = false; uint32_t nb = 0;
    if (t < P.cells.n_cells) {
        Ctx x; locate(P.cells, t, x);
        v = valid_rc(P.cells, x, x.r, x.c);
        if (v) nb = neighbour_mask(P, x);
    }
    const unsigned bv = __ballot_sync(0xffffffffu, v);
    const int pv = __popc(bv & ((1u << lane) - 1u));
    const int deg = __popc(nb);
    int incl = deg;
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) { s_v[warp] = __popc(bv); s_d[warp] = incl; }
    __syncthreads();
    int wv = 0, wd = 0;
    for (int w = 0; w < warp; ++w) { wv += s_v[w]; wd += s_d[w]; }
    const int64_t nb1 = P.nblocks + 1;
    const int id = table[blockIdx.x] + wv + pv;
    const int ep = table[nb1 + blockIdx.x] + wd + (incl - deg);
    if (t < P.cells.n_cells) node_id[t] = v ? id : -1;
    if (v) {
        if (node_cell) node_cell[id] = (int32_t)t;
        if (rowptr) rowptr[id] = ep;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && rowptr) rowptr[table[P.nblocks]] = table[nb1 + P.nblocks];
}

// ---------------------------------------------------------------- node features
struct BoxStats { float mean, stdv; };

// graph_construction.py:399-430: masked 5x5 sums in fp64, zero outside the grid
__device__ BoxStats box_stats(const bgnn_cells& cs, const Ctx& x, int r, int c) {
    double S = 0.0, Q = 0.0; int C = 0;
    for (int dr = -2; dr <= 2; ++dr) {
        const int rr = r + dr;
        if (rr < 0 || rr >= x.H) continue;
        for (int dc = -2; dc <= 2; ++dc) {
            const int cc = c + dc;
            if (cc < 0 || cc >= x.W) continue;
            const int64_t di = x.dbase + (int64_t)rr * x.dstride + cc;
            if (valid_at(cs, di)) {
                const double d = (double)__ldg(cs.depth + di);
                S += d; Q += d * d; ++C;
            }
        }
    }
    const double n = C > 0 ? (double)C : 1.0;
    const double mu = S / n;
    double var = Q / n - mu * mu;
    if (var < 0.0) var = 0.0;                 // np.maximum keeps NaN, so no fmax here
    BoxStats o; o.mean = (float)mu; o.stdv = (float)sqrt(var);
    return o;
}

// depth_filled (:281-282): valid ? depth : local_mean, then nan_to_num
__device__ __forceinline__ float filled_at(const bgnn_cells& cs, const Ctx& x, int r, int c) {
    const int64_t di = x.dbase + (int64_t)r * x.dstride + c;
    float f = valid_at(cs, di) ? __ldg(cs.depth + di) : box_stats(cs, x, r, c).mean;
    return nan_to_num_f32(f);
}

__device__ __forceinline__ void edge_attr_write(const GraphParams& P, float* __restrict__ out, int k,
                                                float dz, double rx, double ry) {
    // tgt - src == offset k (graph_construction.py:348-362)
    const double dx = (double)P.dc[k] * rx, dy = (double)P.dr[k] * ry;
    const double dist = sqrt(dx * dx + dy * dy);
    for (int e = 0; e < P.ED; ++e) {
        float v;
        switch (P.ef[e]) {
            case BGNN_EF_DISTANCE: { double d = dist; if (isnan(d)) d = 0.0; v = (float)d; break; }
            case BGNN_EF_DEPTH_DIFF: v = nan_to_num_f32(dz); break;
            case BGNN_EF_SLOPE: {
                double s = 0.0;
                if (dist > 0.0) s = atan((double)dz / dist) * (180.0 / 3.14159265358979323846);
                if (isnan(s)) s = 0.0;
                v = (float)s; break;
            }
            default: v = 0.0f;
        }
        out[e] = v;
    }
}

__global__ void __launch_bounds__(GB_THREADS) gb_emit(const GraphParams P, const int32_t* __restrict__ node_id,
                                                      const int32_t* __restrict__ rowptr, float* __restrict__ xout,
                                                      float* __restrict__ local_std, int32_t* __restrict__ csr_src,
                                                      float* __restrict__ csr_eattr, float* __restrict__ grid_eattr) {
    const int64_t t = (int64_t)blockIdx.x * GB_THREADS + threadIdx.x;
    if (t >= P.cells.n_cells) return;
    const int id = node_id[t];
    if (id < 0) return;
    const bgnn_cells& cs = P.cells;
    Ctx x; locate(cs, t, x);
    const int r = x.r, c = x.c;
    const int64_t di = x.dbase + (int64_t)r * x.dstride + c;
    const float d = __ldg(cs.depth + di);

    if (xout || local_std) {
        const BoxStats st = box_stats(cs, x, r, c);
        if (local_std) local_std[id] = nan_to_num_f32(st.stdv);
        if (xout) {
            const float f0 = nan_to_num_f32(d);
            const float up = r > 0 ? filled_at(cs, x, r - 1, c) : f0;
            const float dn = r < x.H - 1 ? filled_at(cs, x, r + 1, c) : f0;
            const float lf = c > 0 ? filled_at(cs, x, r, c - 1) : f0;
            const float rt = c < x.W - 1 ? filled_at(cs, x, r, c + 1) : f0;
            // np.gradient, unit spacing (:284)
            float gy = 0.f, gx = 0.f;
            if (x.H >= 2) gy = r == 0 ? dn - f0 : (r == x.H - 1 ? f0 - up : (dn - up) / 2.0f);
            if (x.W >= 2) gx = c == 0 ? rt - f0 : (c == x.W - 1 ? f0 - lf : (rt - lf) / 2.0f);
            const float gm = sqrtf(gx * gx + gy * gy);
            // ndimage.laplace, mode='reflect' (:447): per axis fp64 c*(-2) + (l + r), rounded to fp32
            const double t0 = (double)f0 * -2.0 + ((double)up + (double)dn);
            const double t1 = (double)f0 * -2.0 + ((double)lf + (double)rt);
            float lap = (float)t0 + (float)t1;
            int cnt3 = 0;                      // zero-padded 3x3 valid count (:451-456)
            for (int a = -1; a <= 1; ++a)
                for (int b = -1; b <= 1; ++b) cnt3 += valid_rc(cs, x, r + a, c + b) ? 1 : 0;
            if (cnt3 < 3) lap = 0.0f;
            float* row = xout + (int64_t)id * P.F;
            for (int j = 0; j < P.F; ++j) {
                float v;
                switch (P.nf[j]) {
                    case BGNN_NF_DEPTH: v = d; break;
                    case BGNN_NF_LOCAL_MEAN: v = st.mean; break;
                    case BGNN_NF_LOCAL_STD: v = st.stdv; break;
                    case BGNN_NF_GRAD_X: v = gx; break;
                    case BGNN_NF_GRAD_Y: v = gy; break;
                    case BGNN_NF_GRAD_MAG: v = gm; break;
                    case BGNN_NF_CURVATURE: v = lap; break;
                    default: v = cs.unc ? __ldg(cs.unc + di) : 0.0f;
                }
                row[j] = nan_to_num_f32(v);
            }
        }
    }

    if (grid_eattr && P.K <= 8) {
        // dense per-cell layout (slot k <-> in-edge from i - off_k, zeros when absent), kept in registers and stored as
        // 16-byte vectors; the CSR payload is written on the way when requested
        int ep = rowptr ? rowptr[id] : 0;
        float gev[8 * BGNN_MAX_EDGE_FEATS];
#pragma unroll
        for (int q = 0; q < 8 * BGNN_MAX_EDGE_FEATS; ++q) gev[q] = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k >= P.K) continue;
            const int kk = P.neg[k];
            const int jr = r + P.dr[kk], jc = c + P.dc[kk];
            if (!valid_rc(cs, x, jr, jc)) continue;
            if (csr_src) csr_src[ep] = node_id[x.cbase + (int64_t)jr * x.W + jc];
            const float dj = __ldg(cs.depth + x.dbase + (int64_t)jr * x.dstride + jc);
            float ev[BGNN_MAX_EDGE_FEATS] = {0.f, 0.f, 0.f, 0.f};
            edge_attr_write(P, ev, k, d - dj, x.rx, x.ry);
#pragma unroll
            for (int e = 0; e < BGNN_MAX_EDGE_FEATS; ++e) {
                if (e >= P.ED) continue;
                if (csr_eattr) csr_eattr[(int64_t)ep * P.ED + e] = ev[e];
                if (P.ED == 3) gev[k * 3 + e] = ev[e];                 // the reference's 3 edge features: static indices
                else if (P.ED == 4) gev[k * 4 + e] = ev[e];
                else if (P.ED == 2) gev[k * 2 + e] = ev[e];
                else gev[k + e] = ev[e];
            }
            ++ep;
        }
        float* ge = grid_eattr + t * (int64_t)(P.K * P.ED);
        const int n = P.K * P.ED;                                      // multiple of 4 for K in {4, 8}
        if ((n & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (4 * q < n) reinterpret_cast<float4*>(ge)[q] = make_float4(gev[4 * q], gev[4 * q + 1], gev[4 * q + 2], gev[4 * q + 3]);
        } else {
#pragma unroll
            for (int q = 0; q < 8 * BGNN_MAX_EDGE_FEATS; ++q)
                if (q < n) ge[q] = gev[q];
        }
    } else if (csr_src || csr_eattr) {
        int ep = rowptr[id];
        for (int k = 0; k < P.K; ++k) {          // in-edge (j -> i) with j = i - off_k
            const int kk = P.neg[k];
            const int jr = r + P.dr[kk], jc = c + P.dc[kk];
            if (!valid_rc(cs, x, jr, jc)) continue;
            if (csr_src) csr_src[ep] = node_id[x.cbase + (int64_t)jr * x.W + jc];
            if (csr_eattr) {
                const float dj = __ldg(cs.depth + x.dbase + (int64_t)jr * x.dstride + jc);
                edge_attr_write(P, csr_eattr + (int64_t)ep * P.ED, k, d - dj, x.rx, x.ry);
            }
            ++ep;
        }
    }
}

// ---------------------------------------------------------------- tiled emit (single grid, the reference's 8 offsets)
// Lean grid path: x, local_std and the dense per-cell edge features from a shared-memory tile of the depth grid.
// One CTA per 8 x 32 patch; depth + validity staged once with a halo of 3 (5x5 box of a halo-1 cell that needs the
// local-mean fill); the slope needs one fp64 atan per UNDIRECTED edge: each cell computes it for its 4 forward
// neighbours ((0,1), (1,-1), (1,0), (1,1) = offsets 4..7), the reverse direction is the exact negation
// (atan and the fp32 rounding are odd), so in-edge slot k <= 3 of cell i is -S_i[3-k] and slot k >= 4 is S_j[k-4] of
// the backward neighbour j = i - off_k.  Same arithmetic (and the same fp64 summation order) as gb_emit.
constexpr int TP_H = 8, TP_W = 32, TP_HALO = 3, TP_SH = TP_H + 2 * TP_HALO, TP_SW = TP_W + 2 * TP_HALO;

__global__ void __launch_bounds__(TP_H * TP_W) gb_emit_tile(const GraphParams P, const int32_t* __restrict__ node_id,
                                                            float* __restrict__ xout, float* __restrict__ local_std,
                                                            float* __restrict__ grid_eattr) {
    __shared__ float s_d[TP_SH][TP_SW];            // depth (0 where outside the grid)
    __shared__ unsigned char s_v[TP_SH][TP_SW];    // validity
    __shared__ float s_f[TP_H + 2][TP_W + 2];      // depth_filled on the halo-1 region
    __shared__ __align__(16) float s_s[TP_H + 1][TP_W + 2][4];   // forward slopes (degrees, fp32) of rows -1..TP_H-1, cols -1..TP_W
    const bgnn_cells& cs = P.cells;
    const int H = cs.rows, W = cs.cols;
    const int tid = threadIdx.x;
    const int r0 = blockIdx.y * TP_H, c0 = blockIdx.x * TP_W;

    for (int q = tid; q < TP_SH * TP_SW; q += TP_H * TP_W) {
        const int lr = q / TP_SW, lc = q % TP_SW;
        const int r = r0 - TP_HALO + lr, c = c0 - TP_HALO + lc;
        float d = 0.f; bool v = false;
        if (r >= 0 && r < H && c >= 0 && c < W) {
            const int64_t di = (int64_t)r * cs.row_stride + c;
            d = __ldg(cs.depth + di);
            v = cs.mask ? (__ldg(cs.mask + di) != 0) : (isfinite(d) && (!cs.use_nodata || d != cs.nodata));
        }
        s_d[lr][lc] = d; s_v[lr][lc] = v ? 1 : 0;
    }
    __syncthreads();

    // masked 5x5 box sums in fp64 around local cell (lr, lc) (halo coordinates), zero outside the grid (invalid there)
    auto box = [&](int lr, int lc, float& mean, float& stdv) {
        double S = 0.0, Q = 0.0; int C = 0;
#pragma unroll
        for (int dr = -2; dr <= 2; ++dr)
#pragma unroll
            for (int dc = -2; dc <= 2; ++dc)
                if (s_v[lr + dr][lc + dc]) { const double d = (double)s_d[lr + dr][lc + dc]; S += d; Q += d * d; ++C; }
        const double n = C > 0 ? (double)C : 1.0;
        const double mu = S / n;
        double var = Q / n - mu * mu;
        if (var < 0.0) var = 0.0;
        mean = (float)mu; stdv = (float)sqrt(var);
    };
    // depth_filled on the halo-1 region (cells outside the grid are never read by the stencils below)
    for (int q = tid; q < (TP_H + 2) * (TP_W + 2); q += TP_H * TP_W) {
        const int fr = q / (TP_W + 2), fc = q % (TP_W + 2);
        const int lr = fr + TP_HALO - 1, lc = fc + TP_HALO - 1;
        const int r = r0 - 1 + fr, c = c0 - 1 + fc;
        float f = 0.f;
        if (r >= 0 && r < H && c >= 0 && c < W) {
            if (s_v[lr][lc]) f = s_d[lr][lc];
            else { float m, sd; box(lr, lc, m, sd); f = m; }
            f = nan_to_num_f32(f);
        }
        s_f[fr][fc] = f;
    }
    // forward slopes of rows -1..TP_H-1, cols -1..TP_W: S[f] = degrees(atan((d_F - d_B) / dist_f)) for forward offset f
    const double rx = cs.rx, ry = cs.ry;
    for (int q = tid; q < (TP_H + 1) * (TP_W + 2); q += TP_H * TP_W) {
        const int sr = q / (TP_W + 2), sc = q % (TP_W + 2);
        const int lr = sr + TP_HALO - 1, lc = sc + TP_HALO - 1;
        float out4[4] = {0.f, 0.f, 0.f, 0.f};
        if (s_v[lr][lc]) {
            const float dB = s_d[lr][lc];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int dr = P.dr[4 + f], dc = P.dc[4 + f];
                if (!s_v[lr + dr][lc + dc]) continue;
                const double dx = (double)dc * rx, dy = (double)dr * ry;
                const double dist = sqrt(dx * dx + dy * dy);
                const float dz = s_d[lr + dr][lc + dc] - dB;
                double sl = 0.0;
                if (dist > 0.0) sl = atan((double)dz / dist) * (180.0 / 3.14159265358979323846);
                if (isnan(sl)) sl = 0.0;
                out4[f] = (float)sl;
            }
        }
        *reinterpret_cast<float4*>(&s_s[sr][sc][0]) = make_float4(out4[0], out4[1], out4[2], out4[3]);
    }
    __syncthreads();

    const int pr = tid / TP_W, pc = tid % TP_W;
    const int r = r0 + pr, c = c0 + pc;
    if (r >= H || c >= W) return;
    const int lr = pr + TP_HALO, lc = pc + TP_HALO;
    if (!s_v[lr][lc]) return;
    const int64_t t = (int64_t)r * W + c;
    const int id = node_id[t];
    const float d = s_d[lr][lc];

    if (xout || local_std) {
        float mean, stdv;
        box(lr, lc, mean, stdv);
        if (local_std) local_std[id] = nan_to_num_f32(stdv);
        if (xout) {
            const float f0 = nan_to_num_f32(d);
            const float up = r > 0 ? s_f[pr][pc + 1] : f0;
            const float dn = r < H - 1 ? s_f[pr + 2][pc + 1] : f0;
            const float lf = c > 0 ? s_f[pr + 1][pc] : f0;
            const float rt = c < W - 1 ? s_f[pr + 1][pc + 2] : f0;
            float gy = 0.f, gx = 0.f;
            if (H >= 2) gy = r == 0 ? dn - f0 : (r == H - 1 ? f0 - up : (dn - up) / 2.0f);
            if (W >= 2) gx = c == 0 ? rt - f0 : (c == W - 1 ? f0 - lf : (rt - lf) / 2.0f);
            const float gm = sqrtf(gx * gx + gy * gy);
            const double t0 = (double)f0 * -2.0 + ((double)up + (double)dn);
            const double t1 = (double)f0 * -2.0 + ((double)lf + (double)rt);
            float lap = (float)t0 + (float)t1;
            int cnt3 = 0;
#pragma unroll
            for (int a = -1; a <= 1; ++a)
#pragma unroll
                for (int b = -1; b <= 1; ++b) cnt3 += s_v[lr + a][lc + b];
            if (cnt3 < 3) lap = 0.0f;
            float* row = xout + (int64_t)id * P.F;
            for (int j = 0; j < P.F; ++j) {
                float v;
                switch (P.nf[j]) {
                    case BGNN_NF_DEPTH: v = d; break;
                    case BGNN_NF_LOCAL_MEAN: v = mean; break;
                    case BGNN_NF_LOCAL_STD: v = stdv; break;
                    case BGNN_NF_GRAD_X: v = gx; break;
                    case BGNN_NF_GRAD_Y: v = gy; break;
                    case BGNN_NF_GRAD_MAG: v = gm; break;
                    case BGNN_NF_CURVATURE: v = lap; break;
                    default: v = cs.unc ? __ldg(cs.unc + (int64_t)r * cs.row_stride + c) : 0.0f;
                }
                row[j] = nan_to_num_f32(v);
            }
        }
    }
    if (grid_eattr) {
        float gev[8 * BGNN_MAX_EDGE_FEATS];
#pragma unroll
        for (int q = 0; q < 8 * BGNN_MAX_EDGE_FEATS; ++q) gev[q] = 0.0f;
        const float4 mine = *reinterpret_cast<const float4*>(&s_s[pr + 1][pc + 1][0]);
        const float ms[4] = {mine.x, mine.y, mine.z, mine.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int jr = lr - P.dr[k], jc = lc - P.dc[k];          // source = i - off_k
            if (!s_v[jr][jc]) continue;
            const double dx = (double)P.dc[k] * rx, dy = (double)P.dr[k] * ry;
            double dist = sqrt(dx * dx + dy * dy);
            const float dz = d - s_d[jr][jc];
            const float slope = k < 4 ? 0.0f - ms[3 - k] : s_s[pr + 1 - P.dr[k]][pc + 1 - P.dc[k]][k - 4];   // 0 - x: no -0.0
#pragma unroll
            for (int e = 0; e < BGNN_MAX_EDGE_FEATS; ++e) {
                if (e >= P.ED) continue;
                float v;
                switch (P.ef[e]) {
                    case BGNN_EF_DISTANCE: { double dd = dist; if (isnan(dd)) dd = 0.0; v = (float)dd; break; }
                    case BGNN_EF_DEPTH_DIFF: v = nan_to_num_f32(dz); break;
                    case BGNN_EF_SLOPE: v = slope; break;
                    default: v = 0.0f;
                }
                if (P.ED == 3) gev[k * 3 + e] = v;
                else if (P.ED == 4) gev[k * 4 + e] = v;
                else if (P.ED == 2) gev[k * 2 + e] = v;
                else gev[k + e] = v;
            }
        }
        float* ge = grid_eattr + t * (int64_t)(8 * P.ED);
        const int n = 8 * P.ED;
        if ((n & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (4 * q < n) reinterpret_cast<float4*>(ge)[q] = make_float4(gev[4 * q], gev[4 * q + 1], gev[4 * q + 2], gev[4 * q + 3]);
        } else {
#pragma unroll
            for (int q = 0; q < 8 * BGNN_MAX_EDGE_FEATS; ++q)
                if (q < n) ge[q] = gev[q];
        }
    }
}

// ---------------------------------------------------------------- banded emit (ragged batches, the reference's 8 offsets)
// gb_emit_tile's data flow for a ragged batch of small grids: one CTA per BAND PATCH (a band of rows of one grid over its
// full width, <= 256 cells; the table the host plans for the band-patch aggregation, include/bgnn.h: bgnn_grid_topo::patches).
// The generic gb_emit finds every cell's grid with a 20-step binary search and recomputes 5x5 boxes and eight fp64 atans
// per cell from global memory: 2.3 ms per 8 M-cell chunk; here depth + validity of the band (halo 3) are staged once in shared
// memory, the grid comes with the patch, and each cell computes the four FORWARD slopes only (the reverse ones are their
// exact negations, as in gb_emit_tile).  Same arithmetic and fp64 summation order as gb_emit.
struct BandPatch { long long cell_off; int gh, gw, r0, nr; };
constexpr int BP_CELLS = 1024, BP_INNER = 324, BP_HALO = 3;

__global__ void __launch_bounds__(256) gb_emit_bands(const GraphParams P, const BandPatch* __restrict__ patches,
                                                     const int32_t* __restrict__ node_id, float* __restrict__ xout,
                                                     float* __restrict__ local_std, float* __restrict__ grid_eattr) {
    __shared__ float s_d[BP_CELLS];                 // depth (0 outside the grid), (nr + 6) x (W + 6)
    __shared__ unsigned char s_v[BP_CELLS];         // validity
    __shared__ float s_f[BP_INNER];                 // depth_filled on the halo-1 region, (nr + 2) x (W + 2)
    __shared__ __align__(16) float s_s[BP_INNER][4];   // forward slopes of rows -1 .. nr-1, cols -1 .. W
    const bgnn_cells& cs = P.cells;
    const BandPatch pt = patches[blockIdx.x];
    const int H = pt.gh, W = pt.gw, nr = pt.nr, r0 = pt.r0;
    const int SW = W + 2 * BP_HALO, SH = nr + 2 * BP_HALO, FW = W + 2;
    const int tid = threadIdx.x;
    const int g = (int)0;
    (void)g;
    // per-grid resolution: found from the packed offset (one binary search per CTA, by one thread)
    __shared__ double s_res[2];
    if (tid == 0) {
        int lo = 0, hi = cs.n_grids;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(cs.cell_off + mid) <= pt.cell_off) lo = mid; else hi = mid; }
        s_res[0] = __ldg(cs.grid_res + 2 * lo); s_res[1] = __ldg(cs.grid_res + 2 * lo + 1);
    }
    for (int q = tid; q < SH * SW; q += 256) {
        const int lr = q / SW, lc = q - lr * SW;
        const int r = r0 - BP_HALO + lr, c = lc - BP_HALO;
        float d = 0.f; bool v = false;
        if (r >= 0 && r < H && c >= 0 && c < W) {
            const int64_t di = pt.cell_off + (int64_t)r * W + c;
            d = __ldg(cs.depth + di);
            v = cs.mask ? (__ldg(cs.mask + di) != 0) : (isfinite(d) && (!cs.use_nodata || d != cs.nodata));
        }
        s_d[q] = d; s_v[q] = v ? 1 : 0;
    }
    __syncthreads();
    auto box = [&](int lr, int lc, float& mean, float& stdv) {     // halo coordinates; zero outside the grid (invalid there)
        double S = 0.0, Q = 0.0; int C = 0;
#pragma unroll
        for (int dr = -2; dr <= 2; ++dr)
#pragma unroll
            for (int dc = -2; dc <= 2; ++dc) {
                const int q = (lr + dr) * SW + lc + dc;
                if (s_v[q]) { const double d = (double)s_d[q]; S += d; Q += d * d; ++C; }
            }
        const double n = C > 0 ? (double)C : 1.0;
        const double mu = S / n;
        double var = Q / n - mu * mu;
        if (var < 0.0) var = 0.0;
        mean = (float)mu; stdv = (float)sqrt(var);
    };
    for (int q = tid; q < (nr + 2) * FW; q += 256) {               // depth_filled on the halo-1 region
        const int fr = q / FW, fc = q - fr * FW;
        const int lr = fr + BP_HALO - 1, lc = fc + BP_HALO - 1;
        const int r = r0 - 1 + fr, c = fc - 1;
        float f = 0.f;
        if (r >= 0 && r < H && c >= 0 && c < W) {
            if (s_v[lr * SW + lc]) f = s_d[lr * SW + lc];
            else { float m, sd; box(lr, lc, m, sd); f = m; }
            f = nan_to_num_f32(f);
        }
        s_f[q] = f;
    }
    const double rx = s_res[0], ry = s_res[1];
    for (int q = tid; q < (nr + 1) * FW; q += 256) {               // forward slopes of rows -1 .. nr-1, cols -1 .. W
        const int sr = q / FW, sc = q - sr * FW;
        const int lr = sr + BP_HALO - 1, lc = sc + BP_HALO - 1;
        float out4[4] = {0.f, 0.f, 0.f, 0.f};
        if (s_v[lr * SW + lc]) {
            const float dB = s_d[lr * SW + lc];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int dr = P.dr[4 + f], dc = P.dc[4 + f];
                if (!s_v[(lr + dr) * SW + lc + dc]) continue;
                const double dx = (double)dc * rx, dy = (double)dr * ry;
                const double dist = sqrt(dx * dx + dy * dy);
                const float dz = s_d[(lr + dr) * SW + lc + dc] - dB;
                double sl = 0.0;
                if (dist > 0.0) sl = atan((double)dz / dist) * (180.0 / 3.14159265358979323846);
                if (isnan(sl)) sl = 0.0;
                out4[f] = (float)sl;
            }
        }
        *reinterpret_cast<float4*>(&s_s[q][0]) = make_float4(out4[0], out4[1], out4[2], out4[3]);
    }
    __syncthreads();

    if (tid >= nr * W) return;
    const int pr = tid / W, pc = tid - pr * W;
    const int r = r0 + pr, c = pc;
    const int lr = pr + BP_HALO, lc = pc + BP_HALO;
    if (!s_v[lr * SW + lc]) return;
    const int64_t t = pt.cell_off + (int64_t)r * W + c;
    const int id = node_id[t];
    const float d = s_d[lr * SW + lc];
    if (xout || local_std) {
        float mean, stdv;
        box(lr, lc, mean, stdv);
        if (local_std) local_std[id] = nan_to_num_f32(stdv);
        if (xout) {
            const float f0 = nan_to_num_f32(d);
            const float up = r > 0 ? s_f[pr * FW + pc + 1] : f0;
            const float dn = r < H - 1 ? s_f[(pr + 2) * FW + pc + 1] : f0;
            const float lf = c > 0 ? s_f[(pr + 1) * FW + pc] : f0;
            const float rt = c < W - 1 ? s_f[(pr + 1) * FW + pc + 2] : f0;
            float gy = 0.f, gx = 0.f;
            if (H >= 2) gy = r == 0 ? dn - f0 : (r == H - 1 ? f0 - up : (dn - up) / 2.0f);
            if (W >= 2) gx = c == 0 ? rt - f0 : (c == W - 1 ? f0 - lf : (rt - lf) / 2.0f);
            const float gm = sqrtf(gx * gx + gy * gy);
            const double t0 = (double)f0 * -2.0 + ((double)up + (double)dn);
            const double t1 = (double)f0 * -2.0 + ((double)lf + (double)rt);
            float lap = (float)t0 + (float)t1;
            int cnt3 = 0;
#pragma unroll
            for (int a = -1; a <= 1; ++a)
#pragma unroll
                for (int b = -1; b <= 1; ++b) cnt3 += s_v[(lr + a) * SW + lc + b];
            if (cnt3 < 3) lap = 0.0f;
            float* row = xout + (int64_t)id * P.F;
            for (int j = 0; j < P.F; ++j) {
                float v;
                switch (P.nf[j]) {
                    case BGNN_NF_DEPTH: v = d; break;
                    case BGNN_NF_LOCAL_MEAN: v = mean; break;
                    case BGNN_NF_LOCAL_STD: v = stdv; break;
                    case BGNN_NF_GRAD_X: v = gx; break;
                    case BGNN_NF_GRAD_Y: v = gy; break;
                    case BGNN_NF_GRAD_MAG: v = gm; break;
                    case BGNN_NF_CURVATURE: v = lap; break;
                    default: v = cs.unc ? __ldg(cs.unc + t) : 0.0f;
                }
                row[j] = nan_to_num_f32(v);
            }
        }
    }
    if (grid_eattr) {
        float gev[8 * BGNN_MAX_EDGE_FEATS];
#pragma unroll
        for (int q = 0; q < 8 * BGNN_MAX_EDGE_FEATS; ++q) gev[q] = 0.0f;
        const float4 mine = *reinterpret_cast<const float4*>(&s_s[(pr + 1) * FW + pc + 1][0]);
        const float ms[4] = {mine.x, mine.y, mine.z, mine.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int jr = lr - P.dr[k], jc = lc - P.dc[k];          // source = i - off_k
            if (!s_v[jr * SW + jc]) continue;
            const double dx = (double)P.dc[k] * rx, dy = (double)P.dr[k] * ry;
            double dist = sqrt(dx * dx + dy * dy);
            const float dz = d - s_d[jr * SW + jc];
            const float slope = k < 4 ? 0.0f - ms[3 - k] : s_s[(pr + 1 - P.dr[k]) * FW + pc + 1 - P.dc[k]][k - 4];   // 0 - x: no -0.0
#pragma unroll
            for (int e = 0; e < BGNN_MAX_EDGE_FEATS; ++e) {
                if (e >= P.ED) continue;
                float v;
                switch (P.ef[e]) {
                    case BGNN_EF_DISTANCE: { double dd = dist; if (isnan(dd)) dd = 0.0; v = (float)dd; break; }
                    case BGNN_EF_DEPTH_DIFF: v = nan_to_num_f32(dz); break;
                    case BGNN_EF_SLOPE: v = slope; break;
                    default: v = 0.0f;
                }
                if (P.ED == 3) gev[k * 3 + e] = v;
                else if (P.ED == 4) gev[k * 4 + e] = v;
                else if (P.ED == 2) gev[k * 2 + e] = v;
                else gev[k + e] = v;
            }
        }
        float* ge = grid_eattr + t * (int64_t)(8 * P.ED);
        const int n = 8 * P.ED;
        if ((n & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (4 * q < n) reinterpret_cast<float4*>(ge)[q] = make_float4(gev[4 * q], gev[4 * q + 1], gev[4 * q + 2], gev[4 * q + 3]);
        } else {
#pragma unroll
            for (int q = 0; q < 8 * BGNN_MAX_EDGE_FEATS; ++q)
                if (q < n) ge[q] = gev[q];
        }
    }
}

// reference COO: offset-major, node-minor (graph_construction.py:196-241)
__global__ void __launch_bounds__(GB_THREADS) gb_coo(const GraphParams P, const int32_t* __restrict__ table,
                                                     const int32_t* __restrict__ node_id, int64_t E,
                                                     int64_t* __restrict__ edge_index, float* __restrict__ edge_attr) {
    __shared__ int s_cnt[GB_WARPS][BGNN_MAX_OFFSETS];
    __shared__ int64_t s_goff[BGNN_MAX_OFFSETS];
    const int64_t t = (int64_t)blockIdx.x * GB_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nb1 = P.nblocks + 1;
    const bgnn_cells& cs = P.cells;
    bool v = false; uint32_t nb = 0; Ctx x;
    if (t < cs.n_cells) {
        locate(cs, t, x);
        v = node_id[t] >= 0;
        if (v) nb = neighbour_mask(P, x);
    }
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int k = 0; k < P.K; ++k) { s_goff[k] = run; run += table[(int64_t)(2 + k) * nb1 + P.nblocks]; }
    }
    int rank[BGNN_MAX_OFFSETS];
    for (int k = 0; k < P.K; ++k) {
        const unsigned bk = __ballot_sync(0xffffffffu, (nb >> k) & 1u);
        rank[k] = __popc(bk & ((1u << lane) - 1u));
        if (lane == 0) s_cnt[warp][k] = __popc(bk);
    }
    __syncthreads();
    if (!v) return;
    const int64_t di = x.dbase + (int64_t)x.r * x.dstride + x.c;
    const float d = __ldg(cs.depth + di);
    const int id = node_id[t];
    for (int k = 0; k < P.K; ++k) {
        if (!((nb >> k) & 1u)) continue;
        int wb = 0;
        for (int w = 0; w < warp; ++w) wb += s_cnt[w][k];
        const int64_t e = s_goff[k] + table[(int64_t)(2 + k) * nb1 + blockIdx.x] + wb + rank[k];
        const int jr = x.r + P.dr[k], jc = x.c + P.dc[k];
        if (edge_index) {
            edge_index[e] = id;
            edge_index[E + e] = node_id[x.cbase + (int64_t)jr * x.W + jc];
        }
        if (edge_attr) {
            const float dn = __ldg(cs.depth + x.dbase + (int64_t)jr * x.dstride + jc);
            edge_attr_write(P, edge_attr + e * P.ED, k, dn - d, x.rx, x.ry);
        }
    }
}

int make_params(const bgnn_cells* cells, const bgnn_graph_cfg* cfg, GraphParams& P) {
    BGNN_CHECK_ARG(cells && cfg, "null cells/cfg");
    BGNN_CHECK_ARG(cells->n_grids >= 1 && cells->n_cells >= 0, "bad cell space");
    BGNN_CHECK_ARG(cells->n_cells < (int64_t)2000000000, "cell space too large for int32 node ids");
    BGNN_CHECK_ARG(cells->depth || cells->n_cells == 0, "depth is null");
    BGNN_CHECK_ARG(cfg->n_offsets >= 1 && cfg->n_offsets <= BGNN_MAX_OFFSETS, "n_offsets out of range");
    BGNN_CHECK_ARG(cfg->n_node_feats >= 0 && cfg->n_node_feats <= BGNN_MAX_NODE_FEATS, "n_node_feats out of range");
    BGNN_CHECK_ARG(cfg->n_edge_feats >= 0 && cfg->n_edge_feats <= BGNN_MAX_EDGE_FEATS, "n_edge_feats out of range");
    if (cells->n_grids == 1) {
        BGNN_CHECK_ARG((int64_t)cells->rows * cells->cols == cells->n_cells, "rows*cols != n_cells");
        BGNN_CHECK_ARG(cells->row_stride >= cells->cols, "row_stride < cols");
    } else {
        BGNN_CHECK_ARG(cells->cell_off && cells->grid_rows && cells->grid_cols && cells->grid_res,
                       "ragged descriptors missing");
    }
    P.cells = *cells;
    P.K = cfg->n_offsets;
    for (int k = 0; k < P.K; ++k) { P.dr[k] = cfg->off_dr[k]; P.dc[k] = cfg->off_dc[k]; P.neg[k] = -1; }
    for (int k = 0; k < P.K; ++k)
        for (int j = 0; j < P.K; ++j)
            if (P.dr[j] == -P.dr[k] && P.dc[j] == -P.dc[k]) P.neg[k] = j;
    for (int k = 0; k < P.K; ++k)
        BGNN_CHECK_ARG(P.neg[k] >= 0 && !(P.dr[k] == 0 && P.dc[k] == 0),
                       "offset set must be symmetric and exclude (0,0)");
    P.F = cfg->n_node_feats;
    for (int j = 0; j < P.F; ++j) P.nf[j] = cfg->node_feats[j];
    P.ED = cfg->n_edge_feats;
    for (int j = 0; j < P.ED; ++j) P.ef[j] = cfg->edge_feats[j];
    P.ncol = P.K + 2;
    P.nblocks = ceil_div64(cells->n_cells, GB_THREADS);
    return BGNN_OK;
}

}  // namespace

extern "C" size_t bgnn_graph_scratch_bytes(int64_t n_cells, int32_t n_offsets) {
    const int64_t nblocks = ceil_div64(n_cells < 0 ? 0 : n_cells, GB_THREADS);
    return align_up((size_t)(n_offsets + 2) * (size_t)(nblocks + 1) * sizeof(int32_t), 256) + 256;   // + 64-bit totals
}

static long long* totals64_of(void* scratch, const GraphParams& P) {
    return (long long*)((char*)scratch + align_up((size_t)P.ncol * (size_t)(P.nblocks + 1) * sizeof(int32_t), 256));
}

extern "C" int bgnn_graph_count(const bgnn_cells* cells, const bgnn_graph_cfg* cfg, void* scratch,
                                size_t scratch_bytes, int64_t* totals_host, void* stream_) {
    GraphParams P;
    int rc = make_params(cells, cfg, P);
    if (rc) return rc;
    BGNN_CHECK_ARG(totals_host, "totals_host is null");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (cells->n_cells == 0) { totals_host[0] = totals_host[1] = 0; return BGNN_OK; }
    BGNN_CHECK_ARG(scratch && scratch_bytes >= bgnn_graph_scratch_bytes(cells->n_cells, P.K), "scratch too small");
    int32_t* table = (int32_t*)scratch;
    { ProfScope ps(0, stream); gb_count<<<(unsigned)P.nblocks, GB_THREADS, 0, stream>>>(P, table); }
    BGNN_LAUNCH_CHECK();
    long long* tot64 = totals64_of(scratch, P);
    { ProfScope ps(1, stream); gb_scan<<<P.ncol, 1024, 0, stream>>>(table, P.nblocks, P.ncol, tot64); }
    BGNN_LAUNCH_CHECK();
    long long tot[2];
    BGNN_CUDA(cudaMemcpyAsync(tot, tot64, 2 * sizeof(long long), cudaMemcpyDeviceToHost, stream));
    BGNN_CUDA(cudaStreamSynchronize(stream));
    totals_host[0] = tot[0];
    totals_host[1] = tot[1];
    // node ids, rowptr and the scan table are int32: refuse a space whose edge list does not fit (the totals above
    // are exact, so the caller can still report the size)
    BGNN_CHECK_ARG(tot[0] < (long long)INT32_MAX && tot[1] < (long long)INT32_MAX,
                   "graph too large: node / edge count does not fit int32 CSR offsets (split the cell space)");
    return BGNN_OK;
}

extern "C" int bgnn_graph_emit(const bgnn_cells* cells, const bgnn_graph_cfg* cfg, void* scratch,
                               size_t scratch_bytes, int32_t* node_id, int32_t* node_cell, float* x,
                               float* local_std, int32_t* rowptr, int32_t* csr_src, float* csr_eattr,
                               float* grid_eattr, void* stream_) {
    GraphParams P;
    int rc = make_params(cells, cfg, P);
    if (rc) return rc;
    if (cells->n_cells == 0) return BGNN_OK;
    BGNN_CHECK_ARG(scratch && scratch_bytes >= bgnn_graph_scratch_bytes(cells->n_cells, P.K), "scratch too small");
    BGNN_CHECK_ARG(node_id, "node_id is required");
    BGNN_CHECK_ARG(rowptr || !(csr_src || csr_eattr), "rowptr is required for CSR outputs");
    BGNN_CHECK_ARG(!grid_eattr || P.K <= 8, "grid_eattr needs at most 8 offsets");
    BGNN_CHECK_ARG(!(x && P.F == 0), "x requested with zero node features");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int32_t* table = (const int32_t*)scratch;
    { ProfScope ps(2, stream); gb_ids<<<(unsigned)P.nblocks, GB_THREADS, 0, stream>>>(P, table, node_id, node_cell, rowptr); }
    BGNN_LAUNCH_CHECK();
    if (x || local_std || csr_src || csr_eattr || grid_eattr) {
        ProfScope ps(3, stream);
        // lean grid path (no CSR payload requested, the reference's 8 offsets in their order): shared-memory tiled kernel
        static const int r8[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, c8[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
        bool std8 = P.K == 8 && cells->n_grids == 1 && !csr_src && !csr_eattr && cells->rows >= 2 && cells->cols >= 2;
        for (int k = 0; k < 8 && std8; ++k) std8 = P.dr[k] == r8[k] && P.dc[k] == c8[k];
        if (std8) {
            dim3 grid((unsigned)((cells->cols + TP_W - 1) / TP_W), (unsigned)((cells->rows + TP_H - 1) / TP_H));
            gb_emit_tile<<<grid, TP_H * TP_W, 0, stream>>>(P, node_id, x, local_std, grid_eattr);
        } else {
            gb_emit<<<(unsigned)P.nblocks, GB_THREADS, 0, stream>>>(P, node_id, rowptr, x, local_std, csr_src, csr_eattr, grid_eattr);
        }
        BGNN_LAUNCH_CHECK();
    }
    return BGNN_OK;
}

extern "C" int bgnn_graph_emit_bands(const bgnn_cells* cells, const bgnn_graph_cfg* cfg, void* scratch, size_t scratch_bytes,
                                     const void* patches, int32_t n_patches, int32_t* node_id, int32_t* node_cell, float* x,
                                     float* local_std, int32_t* rowptr, float* grid_eattr, void* stream_) {
    GraphParams P;
    int rc = make_params(cells, cfg, P);
    if (rc) return rc;
    if (cells->n_cells == 0) return BGNN_OK;
    BGNN_CHECK_ARG(scratch && scratch_bytes >= bgnn_graph_scratch_bytes(cells->n_cells, P.K), "scratch too small");
    BGNN_CHECK_ARG(node_id && patches && n_patches > 0 && cells->n_grids > 1, "ragged batch with a band-patch table required");
    static const int r8[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, c8[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    bool std8 = P.K == 8;
    for (int k = 0; k < 8 && std8; ++k) std8 = P.dr[k] == r8[k] && P.dc[k] == c8[k];
    BGNN_CHECK_ARG(std8, "the banded emit covers the reference's 8-connected offset list");
    BGNN_CHECK_ARG(!(x && P.F == 0), "x requested with zero node features");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int32_t* table = (const int32_t*)scratch;
    { ProfScope ps(2, stream); gb_ids<<<(unsigned)P.nblocks, GB_THREADS, 0, stream>>>(P, table, node_id, node_cell, rowptr); }
    BGNN_LAUNCH_CHECK();
    { ProfScope ps(3, stream);
      gb_emit_bands<<<(unsigned)n_patches, 256, 0, stream>>>(P, reinterpret_cast<const BandPatch*>(patches), node_id, x, local_std, grid_eattr); }
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_graph_coo(const bgnn_cells* cells, const bgnn_graph_cfg* cfg, void* scratch,
                              size_t scratch_bytes, const int32_t* node_id, int64_t n_edges,
                              int64_t* edge_index, float* edge_attr, void* stream_) {
    GraphParams P;
    int rc = make_params(cells, cfg, P);
    if (rc) return rc;
    if (cells->n_cells == 0 || n_edges == 0) return BGNN_OK;
    BGNN_CHECK_ARG(scratch && scratch_bytes >= bgnn_graph_scratch_bytes(cells->n_cells, P.K), "scratch too small");
    BGNN_CHECK_ARG(node_id, "node_id is required");
    cudaStream_t stream = (cudaStream_t)stream_;
    { ProfScope ps(4, stream);
      gb_coo<<<(unsigned)P.nblocks, GB_THREADS, 0, stream>>>(P, (const int32_t*)scratch, node_id, n_edges,
                                                            edge_index, edge_attr); }
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}
