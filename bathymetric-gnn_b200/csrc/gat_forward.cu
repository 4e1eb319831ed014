// GAT forward on the device (sm_100a).
//
// Replaces BathymetricGNN.forward / predict of the reference (models/gnn.py:360-451):
//   LocalFeatureExtractor (gnn.py:55-68)
//   -> L x [ PyG GATConv(edge_dim) (gnn.py:125-132, 175-176) -> BatchNorm eval (151-154, 181) -> ReLU (184-186) ]
//   -> classification / confidence / correction heads (gnn.py:191-260, 392-405) -> thresholds (436-449).
// GATConv semantics follow SURVEY.md 3.4 (public PyG definition; PyG itself is not vendored).
//
// Kernels
//   gemm_nt          C = A . W^T (+bias, ReLU) with an optional epilogue that reduces
//                    alpha_src / alpha_dst = sum_c h[n,h,c] * att[h,c] while the tile is in registers
//   gat_aggregate    one warp per destination node: lane-per-edge scores, warp-shuffle segment
//                    softmax, weighted gather of the source rows, +bias, BN(eval), ReLU.  Per-edge
//                    tensors ([E+N, H*C] in PyG) are never materialised; the self-loop edge attribute
//                    (mean of the incoming edge attributes, fill_value='mean') is reduced on the fly.
//   heads_finish     second layer of the three heads + softmax / argmax / sigmoid / action
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include <array>
#include <vector>

#include "common.cuh"
#include "proj_umma.cuh"

// Kernel-selection switches live in the MODEL HANDLE (bgnn_model::opt, bgnn_model_set_option): the library keeps no mutable
// process-wide state, so handles on different devices / host threads never see each other's settings.  The BGNN_NO_*
// environment variables only seed the defaults of a handle at creation (A/B runs of the bench without code changes).
struct ModelOpts {
    bool use_umma = true;           // tensor-core (tcgen05) projections
    bool presplit = true;           // tiled aggregation writes fp16 hi / lo planes for the next tcgen05 projection
    bool window = true;             // ragged batches: windowed shared-memory CSR aggregation
    bool tile5 = true;              // 5x5 stencil / bf16 mode on a single grid: radius-2 tiled aggregation
    bool tile5_window = true;       // full 5x5 stencil: sliding register window variant
    bool ragged_patches = true;     // ragged batches with a band-patch table: tiled aggregation per band of grid rows
};
static ModelOpts default_opts() {
    ModelOpts o;
    o.presplit = getenv("BGNN_NO_PRESPLIT") == nullptr;
    o.window = getenv("BGNN_NO_WINDOW") == nullptr;
    o.tile5 = getenv("BGNN_NO_TILE5") == nullptr;
    o.tile5_window = getenv("BGNN_NO_TILE5_WINDOW") == nullptr;
    o.ragged_patches = getenv("BGNN_NO_RAGGED_PATCHES") == nullptr;
    return o;
}
#ifdef BGNN_AGG_TIMING
extern "C" int bgnn_debug_agg_timing(unsigned long long* out8, int reset);
#endif

struct LayerDev {
    int in, H, C, HC, out;
    float *W, *att_src, *att_dst, *Medge, *bias, *bn_a, *bn_b;
    uint16_t* Wsplit;        // 2 x fp16 split of W for the tcgen05 path (nullptr if the shape is not compiled)
    float w_inv;             // 1 / (power-of-two scale applied to W before the split)
    float out_mul, out_add;  // |layer output| <= out_mul * max|h| + out_add   (BN(eval) after a convex combination + bias)
    // H*C = 512 (wide variant): no single tcgen05 instantiation covers N = 512, so the projection runs as nblk column blocks of
    // 128 (whole heads), each its own launch with its own fp16 split of the block's rows of W
    int nblk;
    uint16_t* Wsplit_blk[4];
    float w_inv_blk[4];
    float* Wr;               // bf16 mode: W rounded to bf16 (RN), kept as fp32 values for the FMA GEMM (nullptr until requested)
    uint16_t* Wbf;           // bf16 mode: the same values as bf16 bit patterns [HC, in] for the tensor-core GEMM
};

struct bgnn_model {
    bgnn_model_desc d;
    int device;
    int nheads_mlp;          // 2 or 3 first-layer head blocks (cls, conf[, corr])
    float* blob;             // single device allocation
    uint16_t* blob16;        // bf16 weight splits
    float *feW0, *feb0, *feW1, *feb1;
    std::vector<LayerDev> layers;
    float *hW0, *hb0;        // concatenated first layers [nheads_mlp*h2, C]
    float *Wcls, *bcls, *Wconf, *bconf, *Wcorr, *bcorr;
    int hc_max;
    // tensor-core fusions (proj_umma.cu): layer-0 projection with the feature extractor folded in, fused heads
    uint16_t* Wfold0;        // fp16 split of W_lin0 . W_fe1 (nullptr if not compiled for this shape)
    float* bias_fold0;       // [HC_0] = W_lin0 . b_fe1
    float fold0_inv;
    float mlp0_mul, mlp0_add;   // |relu(x W0^T + b0)| <= mlp0_mul * max|x| + mlp0_add
    uint16_t* fe1split;      // fp16 split of feature_extractor.mlp.3.weight: the unfused MLP's second Linear on the tensor cores
    float fe1_inv;
    uint16_t* hWsplit;       // fp16 split of the concatenated head first layers (zero padded to hNpad rows)
    float heads_inv;
    int hNpad;
    int precision;           // BGNN_PREC_F32 | BGNN_PREC_BF16 (bgnn_model_set_precision)
    ModelOpts opt;           // kernel-selection switches (bgnn_model_set_option)
    float* blob_bf;          // bf16 mode: rounded projection weights (fp32 values, then bf16 bit patterns)
};

namespace {

// ------------------------------------------------------------------------------------------ GEMM
constexpr int BM = 128, BK = 16, GEMM_THREADS = 256, PAD = 4;

// bf16 mode (bgnn_model_set_precision): IO_F32 = fp32 in / fp32 out, IO_F32_BF16 = fp32 A, C stored as bf16 (RN),
// IO_BF16 = A and C are bf16 in HBM (A / Cout are then uint16 arrays behind the float pointers).  The arithmetic is
// always fp32 FMA on exactly converted operands; alpha comes from the fp32 accumulators.
enum { IO_F32 = 0, IO_F32_BF16 = 1, IO_BF16 = 2, IO_BF16_F32 = 3 };
__device__ __forceinline__ float bf16_bits_to_f32(uint32_t b) { return __uint_as_float(b << 16); }
__device__ __forceinline__ uint32_t f32_to_bf16_bits(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {      // one packed convert, a in the low half
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
}

template <int BN, int CH, bool ALPHA, int IO = IO_F32>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_nt(const float* __restrict__ A, int lda, const float* __restrict__ W, int K, int64_t M, int N,
        float* __restrict__ Cout, int ldc, const float* __restrict__ bias, int relu,
        const float* __restrict__ att_src, const float* __restrict__ att_dst, float* __restrict__ alpha,
        int Htot) {
    constexpr int TN = BN / 16;              // columns per thread (8 or 4)
    constexpr int NG = TN / 4;               // float4 column groups per thread (2 or 1)
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Ws[2][BK][BN + PAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const bool vec_ok = (K % 4 == 0) && (lda % 4 == 0);

    // global->register staging: A tile 128x16 = 512 float4 (2 per thread); W tile BNx16
    constexpr int A_F4 = BM * BK / 4 / GEMM_THREADS;          // 2
    constexpr int W_F4 = (BN * BK / 4 + GEMM_THREADS - 1) / GEMM_THREADS;   // 2 or 1
    float4 ra[A_F4], rw[W_F4];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            const int row = f >> 2, kq = (f & 3) * 4;
            const int64_t gm = m0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gm < M && IO == IO_BF16) {
                const uint16_t* p = reinterpret_cast<const uint16_t*>(A) + gm * lda + k0 + kq;
                if (vec_ok && k0 + kq + 3 < K) {
                    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
                    v = make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u),
                                    __uint_as_float(t.y << 16), __uint_as_float(t.y & 0xffff0000u));
                } else {
                    if (k0 + kq + 0 < K) v.x = bf16_bits_to_f32(__ldg(p + 0));
                    if (k0 + kq + 1 < K) v.y = bf16_bits_to_f32(__ldg(p + 1));
                    if (k0 + kq + 2 < K) v.z = bf16_bits_to_f32(__ldg(p + 2));
                    if (k0 + kq + 3 < K) v.w = bf16_bits_to_f32(__ldg(p + 3));
                }
            } else if (gm < M) {
                const float* p = A + gm * lda + k0 + kq;
                if (vec_ok && k0 + kq + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
                else {
                    if (k0 + kq + 0 < K) v.x = __ldg(p + 0);
                    if (k0 + kq + 1 < K) v.y = __ldg(p + 1);
                    if (k0 + kq + 2 < K) v.z = __ldg(p + 2);
                    if (k0 + kq + 3 < K) v.w = __ldg(p + 3);
                }
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < W_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f < BN * BK / 4) {
                const int row = f >> 2, kq = (f & 3) * 4;
                const int gn = n0 + row;
                if (gn < N) {
                    const float* p = W + (int64_t)gn * K + k0 + kq;
                    if (vec_ok && k0 + kq + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
                    else {
                        if (k0 + kq + 0 < K) v.x = __ldg(p + 0);
                        if (k0 + kq + 1 < K) v.y = __ldg(p + 1);
                        if (k0 + kq + 2 < K) v.z = __ldg(p + 2);
                        if (k0 + kq + 3 < K) v.w = __ldg(p + 3);
                    }
                }
            }
            rw[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            const int row = f >> 2, kq = (f & 3) * 4;
            As[buf][kq + 0][row] = ra[i].x; As[buf][kq + 1][row] = ra[i].y;
            As[buf][kq + 2][row] = ra[i].z; As[buf][kq + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < W_F4; ++i) {
            const int f = tid + i * GEMM_THREADS;
            if (f < BN * BK / 4) {
                const int row = f >> 2, kq = (f & 3) * 4;
                Ws[buf][kq + 0][row] = rw[i].x; Ws[buf][kq + 1][row] = rw[i].y;
                Ws[buf][kq + 2][row] = rw[i].z; Ws[buf][kq + 3][row] = rw[i].w;
            }
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    int cur = 0;
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const float4 bv = *reinterpret_cast<const float4*>(&Ws[cur][k][g * 64 + tx * 4]);
                b[g * 4 + 0] = bv.x; b[g * 4 + 1] = bv.y; b[g * 4 + 2] = bv.z; b[g * 4 + 3] = bv.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(cur ^ 1);
            __syncthreads();
            cur ^= 1;
        }
    }

    // epilogue: thread owns rows m0+ty*8+i, columns n0 + g*64 + tx*4 + j
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gm = m0 + ty * 8 + i;
        if (ALPHA) {
            float ps[NG], pd[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                ps[g] = 0.f; pd[g] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int gn = n0 + g * 64 + tx * 4 + j;
                    if (gn < N) {
                        ps[g] = fmaf(acc[i][g * 4 + j], __ldg(att_src + gn), ps[g]);
                        pd[g] = fmaf(acc[i][g * 4 + j], __ldg(att_dst + gn), pd[g]);
                    }
                }
            }
            if (NG == 2 && CH >= 128) { ps[0] += ps[1]; pd[0] += pd[1]; }
            constexpr int LG = CH >= 64 ? 16 : CH / 4;       // lanes that share a head within a group
#pragma unroll
            for (int o = LG / 2; o; o >>= 1) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    ps[g] += __shfl_xor_sync(0xffffffffu, ps[g], o);
                    pd[g] += __shfl_xor_sync(0xffffffffu, pd[g], o);
                }
            }
            if ((tx % LG) == 0 && gm < M) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    if (g == 1 && CH >= 128) break;
                    const int gn = n0 + g * 64 + tx * 4;
                    if (gn < N) {
                        const int h = gn / CH;
                        alpha[gm * (2 * Htot) + h] = ps[g];
                        alpha[gm * (2 * Htot) + Htot + h] = pd[g];
                    }
                }
            }
        }
        if (gm < M) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int gn = n0 + g * 64 + tx * 4;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = acc[i][g * 4 + j];
                    if (bias && gn + j < N) v[j] += __ldg(bias + gn + j);
                    if (relu) v[j] = fmaxf(v[j], 0.f);
                }
                // projections (ALPHA) store h head-major: [Htot][M][CH], one contiguous plane per head
                const int64_t off = ALPHA ? ((int64_t)(gn / CH) * M + gm) * CH + (gn % CH) : gm * ldc + gn;
                if (IO == IO_F32_BF16 || IO == IO_BF16) {
                    uint16_t* p = reinterpret_cast<uint16_t*>(Cout) + off;
                    if (gn + 3 < N && (ldc % 4 == 0)) *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
                    else
                        for (int j = 0; j < 4; ++j)
                            if (gn + j < N) p[j] = (uint16_t)f32_to_bf16_bits(v[j]);
                } else {
                    float* p = Cout + off;
                    if (gn + 3 < N && (ldc % 4 == 0)) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
                    else
                        for (int j = 0; j < 4; ++j)
                            if (gn + j < N) p[j] = v[j];
                }
            }
        }
    }
}

int launch_gemm(const float* A, int lda, const float* W, int K, int64_t M, int N, float* C, int ldc,
                const float* bias, int relu, int CH, const float* att_src, const float* att_dst,
                float* alpha, int Htot, cudaStream_t s, int io = IO_F32) {
    if (M == 0) return BGNN_OK;
    const bool al = alpha != nullptr;
    dim3 block(GEMM_THREADS);
    if ((io == IO_F32_BF16 && al) || (io == IO_BF16 && !al)) { bgnn_set_error("internal: bf16 GEMM variant not compiled"); return BGNN_E_UNSUPPORTED; }
#define GO(BN_, CH_, AL_)                                                                              \
    do {                                                                                               \
        dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)((N + BN_ - 1) / BN_));                       \
        if (io == IO_F32)                                                                              \
            gemm_nt<BN_, CH_, AL_, IO_F32><<<grid, block, 0, s>>>(A, lda, W, K, M, N, C, ldc, bias, relu, att_src, \
                                                                  att_dst, alpha, Htot);               \
        else                                                                                           \
            gemm_nt<BN_, CH_, AL_, (AL_ ? IO_BF16 : IO_F32_BF16)><<<grid, block, 0, s>>>(A, lda, W, K, M, N, C, ldc, bias, relu, att_src, \
                                                                  att_dst, alpha, Htot);               \
    } while (0)
    if (N > 64) {
        if (!al) GO(128, 64, false);
        else if (CH == 32) GO(128, 32, true);
        else if (CH == 64) GO(128, 64, true);
        else if (CH == 128) GO(128, 128, true);
        else { bgnn_set_error("unsupported hidden width %d", CH); return BGNN_E_UNSUPPORTED; }
    } else {
        if (!al) GO(64, 64, false);
        else if (CH == 32) GO(64, 32, true);
        else if (CH == 64) GO(64, 64, true);
        else { bgnn_set_error("unsupported hidden width %d for a <=64-wide projection", CH); return BGNN_E_UNSUPPORTED; }
    }
#undef GO
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

// ------------------------------------------------------------------------------------------ aggregation
constexpr int AGG_WARPS = 8;

template <int VEC> struct VecIO;
template <> struct VecIO<1> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void st(float* p, const float* v) { p[0] = v[0]; }
};
template <> struct VecIO<2> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <> struct VecIO<4> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct VecIO<8> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { VecIO<4>::ld(p, v); VecIO<4>::ld(p + 4, v + 4); }
    static __device__ __forceinline__ void st(float* p, const float* v) { VecIO<4>::st(p, v); VecIO<4>::st(p + 4, v + 4); }
};
template <> struct VecIO<16> {
    static __device__ __forceinline__ void ld(const float* p, float* v) { VecIO<8>::ld(p, v); VecIO<8>::ld(p + 8, v + 8); }
    static __device__ __forceinline__ void st(float* p, const float* v) { VecIO<8>::st(p, v); VecIO<8>::st(p + 8, v + 8); }
};

// bf16 rows (bf16 mode): VEC consecutive bf16 values <-> fp32 registers (exact widening, RN narrowing)
template <int VEC> struct VecBF {
    static __device__ __forceinline__ void ld(const uint16_t* p, float* v) {
        if constexpr (VEC == 1) v[0] = bf16_bits_to_f32(__ldg(p));
        else if constexpr (VEC == 2) { const uint32_t t = __ldg(reinterpret_cast<const uint32_t*>(p)); v[0] = __uint_as_float(t << 16); v[1] = __uint_as_float(t & 0xffff0000u); }
        else if constexpr (VEC == 4) { const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
            v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u); v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u); }
        else {
#pragma unroll
            for (int q = 0; q < VEC / 8; ++q) {
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(p) + q);
                v[8 * q + 0] = __uint_as_float(t.x << 16); v[8 * q + 1] = __uint_as_float(t.x & 0xffff0000u);
                v[8 * q + 2] = __uint_as_float(t.y << 16); v[8 * q + 3] = __uint_as_float(t.y & 0xffff0000u);
                v[8 * q + 4] = __uint_as_float(t.z << 16); v[8 * q + 5] = __uint_as_float(t.z & 0xffff0000u);
                v[8 * q + 6] = __uint_as_float(t.w << 16); v[8 * q + 7] = __uint_as_float(t.w & 0xffff0000u);
            }
        }
    }
    static __device__ __forceinline__ void st(uint16_t* p, const float* v) {
        if constexpr (VEC == 1) p[0] = (uint16_t)f32_to_bf16_bits(v[0]);
        else if constexpr (VEC == 2) *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(v[0], v[1]);
        else if constexpr (VEC == 4) *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
        else {
#pragma unroll
            for (int q = 0; q < VEC / 8; ++q)
                reinterpret_cast<uint4*>(p)[q] = make_uint4(pack_bf16x2(v[8 * q + 0], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                                            pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
        }
    }
};

// One warp per destination node i.  In-edges e = rowptr[i] .. rowptr[i+1]-1 (reference edge
// order) plus the self loop as the LAST entry (PyG add_self_loops appends the loops).
// IO (bf16 mode): IO_BF16 = h and the output are bf16 rows, IO_BF16_F32 = bf16 h, fp32 output (last layer -> heads).
template <int H, int VEC, int IO = IO_F32>
__global__ void __launch_bounds__(AGG_WARPS * 32)
gat_aggregate(const float* __restrict__ h, const float* __restrict__ alpha, const int32_t* __restrict__ rowptr,
              const int32_t* __restrict__ csr_src, const float* __restrict__ eattr, int ED,
              const float* __restrict__ Medge, const float* __restrict__ bias, const float* __restrict__ bn_a,
              const float* __restrict__ bn_b, int relu, float slope, float* __restrict__ out, int64_t N) {
    constexpr int HC = 32 * VEC;
    constexpr int C = HC / H;
    __shared__ float s_a[AGG_WARPS][32][H];     // attention weights of the current chunk
    __shared__ int s_j[AGG_WARPS][32];          // source ids of the current chunk
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int my_head = (lane * VEC) / C;
    float M[BGNN_MAX_EDGE_FEATS][H];
#pragma unroll
    for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d)
#pragma unroll
        for (int hh = 0; hh < H; ++hh) M[d][hh] = d < ED ? __ldg(Medge + d * H + hh) : 0.f;

    for (int64_t i = (int64_t)blockIdx.x * AGG_WARPS + warp; i < N; i += (int64_t)gridDim.x * AGG_WARPS) {
        const int start = __ldg(rowptr + i);
        const int deg = __ldg(rowptr + i + 1) - start;
        const int tot = deg + 1;                                  // + self loop
        // self-loop attribute = mean of the incoming edge attributes (0 when deg == 0)
        float lsum[BGNN_MAX_EDGE_FEATS];
#pragma unroll
        for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) lsum[d] = 0.f;
        if (ED > 0) {
            for (int e = lane; e < deg; e += 32)
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d)
                    if (d < ED) lsum[d] += __ldg(eattr + (int64_t)(start + e) * ED + d);
#pragma unroll
            for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) {
                for (int o = 16; o; o >>= 1) lsum[d] += __shfl_xor_sync(0xffffffffu, lsum[d], o);
                lsum[d] = lsum[d] / (float)(deg > 0 ? deg : 1);
            }
        }
        float ad[H];
#pragma unroll
        for (int hh = 0; hh < H; ++hh) ad[hh] = __ldg(alpha + i * (2 * H) + H + hh);

        // score of virtual edge e (0..tot-1) for all heads, computed by the calling lane
        auto score = [&](int e, int& j, float* s) {
            float ea[BGNN_MAX_EDGE_FEATS];
            if (e < deg) {
                j = __ldg(csr_src + start + e);
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ea[d] = d < ED ? __ldg(eattr + (int64_t)(start + e) * ED + d) : 0.f;
            } else {
                j = (int)i;
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ea[d] = lsum[d];
            }
#pragma unroll
            for (int hh = 0; hh < H; ++hh) {
                float ae = 0.f;
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ae = fmaf(ea[d], M[d][hh], ae);
                float v = (__ldg(alpha + (int64_t)j * (2 * H) + hh) + ad[hh]) + ae;
                s[hh] = v > 0.f ? v : v * slope;
            }
        };

        // pass 1: per-head max
        float mx[H];
#pragma unroll
        for (int hh = 0; hh < H; ++hh) mx[hh] = -CUDART_INF_F;
        for (int e0 = 0; e0 < tot; e0 += 32) {
            const int e = e0 + lane;
            if (e < tot) {
                int j; float s[H];
                score(e, j, s);
#pragma unroll
                for (int hh = 0; hh < H; ++hh) mx[hh] = fmaxf(mx[hh], s[hh]);
            }
        }
#pragma unroll
        for (int hh = 0; hh < H; ++hh)
            for (int o = 16; o; o >>= 1) mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], o));
        // pass 2: per-head sum of exp (single-chunk nodes keep their exps for pass 3)
        float sm[H], pe[H];
        int jmine = 0;
#pragma unroll
        for (int hh = 0; hh < H; ++hh) { sm[hh] = 0.f; pe[hh] = 0.f; }
        for (int e0 = 0; e0 < tot; e0 += 32) {
            const int e = e0 + lane;
            if (e < tot) {
                float s[H];
                score(e, jmine, s);
#pragma unroll
                for (int hh = 0; hh < H; ++hh) { pe[hh] = expf(s[hh] - mx[hh]); sm[hh] += pe[hh]; }
            }
        }
#pragma unroll
        for (int hh = 0; hh < H; ++hh) {
            for (int o = 16; o; o >>= 1) sm[hh] += __shfl_xor_sync(0xffffffffu, sm[hh], o);
            sm[hh] += 1e-16f;
        }
        // pass 3: aggregate
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        for (int e0 = 0; e0 < tot; e0 += 32) {
            const int e = e0 + lane;
            __syncwarp();
            if (e < tot) {
                if (tot > 32) {                       // multi-chunk: recompute this chunk's exps
                    float s[H];
                    score(e, jmine, s);
#pragma unroll
                    for (int hh = 0; hh < H; ++hh) pe[hh] = expf(s[hh] - mx[hh]);
                }
                s_j[warp][lane] = jmine;
#pragma unroll
                for (int hh = 0; hh < H; ++hh) s_a[warp][lane][hh] = pe[hh] / sm[hh];
            }
            __syncwarp();
            const int cnt = min(32, tot - e0);
            for (int q = 0; q < cnt; ++q) {
                const int j = s_j[warp][q];
                const float a = s_a[warp][q][my_head];
                float hv[VEC];
                if constexpr (IO == IO_F32) VecIO<VEC>::ld(h + ((int64_t)my_head * N + j) * C + (lane * VEC) % C, hv);       // h is head-major [H][N][C]
                else VecBF<VEC>::ld(reinterpret_cast<const uint16_t*>(h) + ((int64_t)my_head * N + j) * C + (lane * VEC) % C, hv);
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = fmaf(a, hv[v], acc[v]);
            }
        }
        // epilogue: + bias, BN(eval) as x*a + b, ReLU
        float o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int ch = lane * VEC + v;
            float y = acc[v] + __ldg(bias + ch);
            y = fmaf(y, __ldg(bn_a + ch), __ldg(bn_b + ch));
            o[v] = relu ? fmaxf(y, 0.f) : y;
        }
        if constexpr (IO == IO_BF16) VecBF<VEC>::st(reinterpret_cast<uint16_t*>(out) + i * HC + lane * VEC, o);
        else VecIO<VEC>::st(out + i * HC + lane * VEC, o);
    }
}

template <int H, int VEC>
int launch_agg_t(const LayerDev& L, const float* h, const float* alpha, const int32_t* rowptr, const int32_t* src,
                 const float* eattr, int ED, int relu, float slope, float* out, int64_t N, int sms, cudaStream_t s, int io) {
    int64_t blocks = ceil_div64(N, AGG_WARPS);
    const int64_t cap = (int64_t)sms * 8 * 4;            // persistent grid: multiple of the SM count
    if (blocks > cap) blocks = cap;
    if (io == IO_BF16)
        gat_aggregate<H, VEC, IO_BF16><<<(unsigned)blocks, AGG_WARPS * 32, 0, s>>>(h, alpha, rowptr, src, eattr, ED, L.Medge, L.bias,
                                                                                  L.bn_a, L.bn_b, relu, slope, out, N);
    else if (io == IO_BF16_F32)
        gat_aggregate<H, VEC, IO_BF16_F32><<<(unsigned)blocks, AGG_WARPS * 32, 0, s>>>(h, alpha, rowptr, src, eattr, ED, L.Medge, L.bias,
                                                                                      L.bn_a, L.bn_b, relu, slope, out, N);
    else
        gat_aggregate<H, VEC, IO_F32><<<(unsigned)blocks, AGG_WARPS * 32, 0, s>>>(h, alpha, rowptr, src, eattr, ED, L.Medge, L.bias,
                                                                                 L.bn_a, L.bn_b, relu, slope, out, N);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

int launch_agg(const LayerDev& L, const float* h, const float* alpha, const int32_t* rowptr, const int32_t* src,
               const float* eattr, int ED, int relu, float slope, float* out, int64_t N, int sms, cudaStream_t s,
               int io = IO_F32) {
    if (N == 0) return BGNN_OK;
    const int vec = L.HC / 32;
#define CASE(H_, V_) if (L.H == H_ && vec == V_) return launch_agg_t<H_, V_>(L, h, alpha, rowptr, src, eattr, ED, relu, slope, out, N, sms, s, io)
    CASE(1, 1); CASE(1, 2); CASE(1, 4);
    CASE(2, 2); CASE(2, 4); CASE(2, 8);
    CASE(4, 4); CASE(4, 8); CASE(4, 16);
    CASE(8, 8); CASE(8, 16);
#undef CASE
    bgnn_set_error("unsupported attention shape heads=%d hidden=%d", L.H, L.C);
    return BGNN_E_UNSUPPORTED;
}


// ------------------------------------------------------------------------------------------ tiled aggregation
// Grid path: one CTA per PH x 16 patch of cells.  Every source row of the (PH+2) x 18 halo is staged ONCE per head in
// shared memory by a bulk async copy (cp.async.bulk, one 4*C-byte copy per halo cell, completion on an mbarrier);
// destinations then read their neighbour rows from shared memory.  Thread t owns destination cell t for the score /
// softmax phase (edge features and the self-loop mean stay in its registers across the head loop) and leaves the 3x3
// attention weights of its cell in shared memory, indexed by source POSITION; warp w then walks its two patch rows with
// lane <-> channels and a 3x3 sliding register window, so each step loads only the 3 new source rows of the window's
// right column.  The FMAs run in the reference's edge order (offset k ascending, source = i - off_k, self loop last),
// i.e. the same arithmetic as gat_aggregate; absent neighbours contribute fma(0, 0, acc) == acc.
#ifdef BGNN_AGG_TIMING
__device__ unsigned long long g_agg_t[8];
#define AGG_T0() long long t_prev = clock64()
#define AGG_TICK(slot) do { if (threadIdx.x == 0) { long long t_now = clock64(); atomicAdd(&g_agg_t[slot], (unsigned long long)(t_now - t_prev)); t_prev = t_now; } } while (0)
#else
#define AGG_T0()
#define AGG_TICK(slot)
#endif

struct TopoDev {
    int rows, cols;
    const int32_t* node_id;
    int K;
    int dr[8], dc[8];
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int H, int C, int PH> struct TileCfg {
    static constexpr int PW = 16, HR = PH + 2, HW = PW + 2, NSLOT = HR * HW, NT = PH * PW;
    // C == 64: a HALF-WARP per destination (lane <-> 4 channels), the two halves of a warp walk two patch rows at once and
    // multiply with packed fp32 pairs (fma.rn.f32x2): the attention weights are stored DUPLICATED (w, w) so that a pair is one
    // 64-bit operand; 20 floats per destination, patch rows 8 floats apart modulo the 128-byte bank period (the two halves of a
    // warp read two different destinations with one LDS.128: no bank conflict)
#ifndef BGNN_AGG_PAIR_MINH
#define BGNN_AGG_PAIR_MINH 2          /* smallest H that uses the half-warp / packed-pair aggregate phase (H = 1, 8 x 16 patches: 0.62 vs 0.69 ms without) */
#endif
    static constexpr bool PAIR = C == 64 && H >= BGNN_AGG_PAIR_MINH;
    static constexpr int AST = PAIR ? 20 : 12, AROW = PW * AST + (PAIR ? 8 : 0), ASZ = PH * AROW;
    // alpha_src / alpha_dst of the halo arrive by bulk copy (one per run) when a node's 2H floats are a 16 B multiple and
    // the landing area fits the (not yet used) attention-weight buffer; otherwise they are loaded cell by cell
    static constexpr bool ALPHA_BULK = (2 * H * 4) % 16 == 0 && NSLOT * 2 * H <= ASZ;
    static constexpr int RT = 2 * (HW / 2) + 2;      // run table per halo row: count, then (start << 16 | length, node id) pairs
    static constexpr size_t SMEM = (size_t)(NSLOT * C + NSLOT + NSLOT * H + ASZ + BGNN_MAX_EDGE_FEATS * H + 3 * H * C + HR * RT) * 4 + 32;
};

// window position p = dy*3 + dx (dy, dx in 0..2 <-> source at row r-1+dy, column c-1+dx) visited in reference edge order
template <bool CONN4> struct EdgeOrder {
    static constexpr int N = CONN4 ? 5 : 9;
    __host__ __device__ static constexpr int pos(int q) {
        return CONN4 ? (q == 0 ? 7 : q == 1 ? 1 : q == 2 ? 5 : q == 3 ? 3 : 4)
                     : (q == 0 ? 8 : q == 1 ? 7 : q == 2 ? 6 : q == 3 ? 5 : q == 4 ? 3 : q == 5 ? 2 : q == 6 ? 1 : q == 7 ? 0 : 4);
    }
};

__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// packed fp32 pairs (sm_100: two IEEE fp32 operations per issue slot)
__device__ __forceinline__ unsigned long long pk2(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// Optional output format: instead of fp32 rows, the layer output is written as fp16 rows [N][2][H*C] = hi | lo of
// y * sA that the next layer's tensor-core projection consumes directly (proj_umma.cu, PM_SPLIT): same 4 bytes per element,
// same split arithmetic and the same power-of-two scale as the projection's own converter warps would apply (bit-identical
// products), but the projection then needs no conversion stage at all.
struct SplitOut {
    uint16_t* hi;                  // nullptr: fp32 output; else row n = hi[H*C] | lo[H*C]  (lo of a row at + H*C elements)
    const float* amax_src;         // max|h_l| (device); |y| <= mul * max|h_l| + add
    float mul, add;
};

// power-of-two operand scale of the tensor-core GEMM (same rule as proj_umma_kernel): |y * sA| < 1
__device__ __forceinline__ float operand_scale(const float* amax_src, float mul, float add) {
    float bound = add;
    if (amax_src) bound = fmaf(mul, __ldg(amax_src), add);
    const int e = (int)((__float_as_uint(bound) >> 23) & 0xffu);
    const int es = 253 - e;
    return (e > 0 && e < 255 && es >= 1 && es <= 253) ? __uint_as_float((uint32_t)es << 23) : 1.f;
}

template <int H, int C, int PH, bool CONN4>
__global__ void __launch_bounds__(PH * 16, (C <= 64 ? (PH == 8 ? 3 : 2) : 1))
gat_aggregate_tile(const float* __restrict__ h, const float* __restrict__ alpha, const TopoDev T,
                   const float* __restrict__ grid_eattr, int ED,
                   const float* __restrict__ Medge, const float* __restrict__ bias, const float* __restrict__ bn_a,
                   const float* __restrict__ bn_b, int relu, float slope, float* __restrict__ out, int64_t n_nodes,
                   const SplitOut SO) {
    using G = TileCfg<H, C, PH>;
    using O = EdgeOrder<CONN4>;
    constexpr int HW = G::HW, NSLOT = G::NSLOT, NT = G::NT, AST = G::AST, HR = G::HR;
    constexpr int VEC = C / 32, HC = H * C;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* rows = reinterpret_cast<float*>(smem_raw);          // [NSLOT][C]   source rows of the current head
    int* ids = reinterpret_cast<int*>(rows + NSLOT * C);       // [NSLOT]      node id per halo cell (-1 invalid)
    float* asrc = reinterpret_cast<float*>(ids + NSLOT);       // [H][NSLOT]   alpha_src per halo cell
    float* a_sm = asrc + NSLOT * H;                            // [PH][AROW]   3x3 attention weights of the current head (TileCfg)
    float* al_sm = a_sm;                                       // [NSLOT][2H]  landing area of the alpha bulk copies (setup only)
    float* M_s = a_sm + G::ASZ;                                // [ED_MAX][H]  folded edge projection
    float* s_bn = M_s + BGNN_MAX_EDGE_FEATS * H;               // [3][HC]      bias | bn scale | bn shift
    int* run_tab = reinterpret_cast<int*>(s_bn + 3 * HC);      // [HR][RT]     runs of valid cells per halo row
    uint64_t* bar = reinterpret_cast<uint64_t*>(run_tab + HR * G::RT);   // [0] source rows, [1] alpha

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pr0 = blockIdx.y * PH, pc0 = blockIdx.x * 16;
    AGG_T0();
    const float sA = SO.hi ? operand_scale(SO.amax_src, SO.mul, SO.add) : 1.f;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(NT));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar + 1)), "r"(HR));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int s = tid; s < NSLOT; s += NT) {
        const int r = pr0 - 1 + s / HW, c = pc0 - 1 + s % HW;
        int id = -1;
        if (r >= 0 && r < T.rows && c >= 0 && c < T.cols) id = __ldg(T.node_id + (int64_t)r * T.cols + c);
        ids[s] = id;
    }
    for (int i = tid; i < HC; i += NT) { s_bn[i] = __ldg(bias + i); s_bn[HC + i] = __ldg(bn_a + i); s_bn[2 * HC + i] = __ldg(bn_b + i); }
    if (tid < BGNN_MAX_EDGE_FEATS * H) M_s[tid] = (tid / H) < ED ? __ldg(Medge + tid) : 0.f;
    // this destination's dense in-edge features (slot k <-> source i - off_k; zeros where the edge is absent): fetched
    // straight from the per-cell layout, independent of the node ids, so the loads fly during the id round trip
    const int cs = (tid / 16 + 1) * HW + (tid % 16 + 1);
    float ea[8][BGNN_MAX_EDGE_FEATS];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ea[k][d] = 0.f;
    {
        const int r = pr0 + tid / 16, c = pc0 + tid % 16;
        if (ED > 0 && r < T.rows && c < T.cols) {
            const float* ge = grid_eattr + ((int64_t)r * T.cols + c) * (T.K * ED);
            if (T.K == 8 && ED == 3) {
                float v[24];
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const float4 t4 = __ldg(reinterpret_cast<const float4*>(ge) + q);
                    v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int d = 0; d < 3; ++d) ea[k][d] = v[k * 3 + d];
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d)
                        if (k < T.K && d < ED) ea[k][d] = __ldg(ge + k * ED + d);
            }
        }
    }
    __syncthreads();
    AGG_TICK(0);

    // ---- bulk copies: every run of valid cells in a halo row is ONE contiguous chunk, because adjacent valid cells of a
    //      grid row have consecutive node ids, h is head-major ([H][N][C]) and alpha is [N][2H]
    // The runs of a halo row are the same for alpha and for every head: thread t < HR finds those of row t ONCE (the
    // 18-cell scan is a serial chain on the critical path of warp 0) and replays its own table afterwards.
    if (tid < HR) {
        const int s0 = tid * HW;
        int run = -1, n = 0;
        for (int c = 0; c <= HW; ++c) {
            const int id = c < HW ? ids[s0 + c] : -1;
            if (id >= 0 && run < 0) run = c;
            if (id < 0 && run >= 0) {
                run_tab[tid * G::RT + 1 + 2 * n] = (run << 16) | (c - run);
                run_tab[tid * G::RT + 2 + 2 * n] = ids[s0 + run];
                ++n;
                run = -1;
            }
        }
        run_tab[tid * G::RT] = n;
    }
    auto issue_rows = [&](int hh) {
        if (tid < HR) {
            const float* hp = h + (int64_t)hh * n_nodes * C;
            const int s0 = tid * HW, n = run_tab[tid * G::RT];
            for (int i = 0; i < n; ++i) {
                const int sl = run_tab[tid * G::RT + 1 + 2 * i], id = run_tab[tid * G::RT + 2 + 2 * i];
                bulk_g2s(rows + (s0 + (sl >> 16)) * C, hp + (int64_t)id * C, (uint32_t)(sl & 0xffff) * C * 4, bar);
            }
        }
    };
    if (G::ALPHA_BULK && tid < HR) {
        const int s0 = tid * HW, n = run_tab[tid * G::RT];
        for (int i = 0; i < n; ++i) {
            const int sl = run_tab[tid * G::RT + 1 + 2 * i], id = run_tab[tid * G::RT + 2 + 2 * i];
            bulk_g2s(al_sm + (s0 + (sl >> 16)) * 2 * H, alpha + (int64_t)id * 2 * H, (uint32_t)(sl & 0xffff) * 2 * H * 4, bar + 1);
        }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr(bar + 1)) : "memory");
    }
    issue_rows(0);

    // ---- per-destination state (thread t <-> cell t of the patch)
    const int myid = ids[cs];
    float lsum[BGNN_MAX_EDGE_FEATS], ad[H];
    unsigned nbv = 0;
#pragma unroll
    for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) lsum[d] = 0.f;
#pragma unroll
    for (int hh = 0; hh < H; ++hh) ad[hh] = 0.f;
    if (myid >= 0) {
        int deg = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < T.K && ids[cs - T.dr[k] * HW - T.dc[k]] >= 0) {      // in-edge from i - off_k
                nbv |= 1u << k;
                ++deg;
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) lsum[d] += ea[k][d];
            }
#pragma unroll
        for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) lsum[d] = lsum[d] / (float)(deg > 0 ? deg : 1);
    }
    // rows of invalid halo cells stay zero for every head (the bulk copies never touch them)
    for (int s = tid; s < NSLOT; s += NT)
        if (ids[s] < 0) {
#pragma unroll
            for (int q = 0; q < C / 4; ++q) *reinterpret_cast<float4*>(rows + s * C + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    if (G::ALPHA_BULK) {
        mbar_wait_parity(bar + 1, 0);
        for (int s = tid; s < NSLOT; s += NT)
            if (ids[s] >= 0) {
#pragma unroll
                for (int hh = 0; hh < H; ++hh) asrc[hh * NSLOT + s] = al_sm[s * 2 * H + hh];
            }
        if (myid >= 0) {
#pragma unroll
            for (int hh = 0; hh < H; ++hh) ad[hh] = al_sm[cs * 2 * H + H + hh];
        }
    } else {
        for (int s = tid; s < NSLOT; s += NT) {
            const int id = ids[s];
            if (id >= 0) {
#pragma unroll
                for (int hh = 0; hh < H; ++hh) asrc[hh * NSLOT + s] = __ldg(alpha + (int64_t)id * (2 * H) + hh);
            }
        }
        if (myid >= 0) {
#pragma unroll
            for (int hh = 0; hh < H; ++hh) ad[hh] = __ldg(alpha + (int64_t)myid * (2 * H) + H + hh);
        }
    }
    __syncthreads();                 // asrc complete; the alpha landing area (aliases a_sm) is free
    AGG_TICK(1);
    // window position of in-edge k (source = i - off_k)
    int wpos[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) wpos[k] = (1 - T.dr[k]) * 3 + (1 - T.dc[k]);

#pragma unroll 1
    for (int hh = 0; hh < H; ++hh) {
        // 1. stage the source rows of head hh (head 0 was issued during the setup)
        if (hh > 0) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");       // generic accesses of `rows` before async writes
            issue_rows(hh);
        }
        // 2. scores + softmax of this head while the copies are in flight -> a_sm[t][window position]
        float* aw_t = a_sm + (tid / 16) * G::AROW + (tid % 16) * AST;           // this destination's weights
        {
            float4* az = reinterpret_cast<float4*>(aw_t);
#pragma unroll
            for (int q = 0; q < AST / 4; ++q) az[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (myid >= 0) {
            float Mh[BGNN_MAX_EDGE_FEATS], adh = ad[0];
#pragma unroll
            for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) Mh[d] = M_s[d * H + hh];
#pragma unroll
            for (int q = 1; q < H; ++q) adh = (q == hh) ? ad[q] : adh;
            float sc[9], mx = -CUDART_INF_F;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                sc[k] = 0.f;
                if ((nbv >> k) & 1u) {
                    float ae = 0.f;
#pragma unroll
                    for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ae = fmaf(ea[k][d], Mh[d], ae);
                    float v = (asrc[hh * NSLOT + cs - T.dr[k] * HW - T.dc[k]] + adh) + ae;
                    v = v > 0.f ? v : v * slope;
                    sc[k] = v;
                    mx = fmaxf(mx, v);
                }
            }
            {
                float ae = 0.f;
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ae = fmaf(lsum[d], Mh[d], ae);
                float v = (asrc[hh * NSLOT + cs] + adh) + ae;
                v = v > 0.f ? v : v * slope;
                sc[8] = v;
                mx = fmaxf(mx, v);
            }
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((nbv >> k) & 1u) { sc[k] = expf(sc[k] - mx); sum += sc[k]; }
            sc[8] = expf(sc[8] - mx);
            sum += sc[8];
            sum += 1e-16f;
            const float inv = 1.0f / sum;            // one IEEE division, nine multiplies (<= 1 ulp from exp / sum)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((nbv >> k) & 1u) {
                    const float w = sc[k] * inv;
                    if (G::PAIR) *reinterpret_cast<float2*>(aw_t + 2 * wpos[k]) = make_float2(w, w);
                    else aw_t[wpos[k]] = w;
                }
            {
                const float w = sc[8] * inv;
                if (G::PAIR) *reinterpret_cast<float2*>(aw_t + 8) = make_float2(w, w);
                else aw_t[4] = w;
            }
        }
        AGG_TICK(2);
        // 3. all threads arrive (release: a_sm, zero rows), then wait for the arrivals and the copied bytes
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr(bar)) : "memory");
        mbar_wait_parity(bar, (uint32_t)(hh & 1));
        AGG_TICK(3);
        if constexpr (G::PAIR) {
        // 4. aggregate, C == 64: half-warp <-> destination (lane <-> 4 channels); the two halves of warp w walk patch rows 2w
        //    and 2w + 1 together; 3x3 sliding window of float4 in registers (3 new LDS.128 per step); FMAs on packed fp32 pairs
        //    (two IEEE fmas per issue slot) in the reference's edge order; epilogue on pairs as well
        const int l16 = lane & 15, pr = warp * 2 + (lane >> 4);
        const int ch0 = hh * C + l16 * 4;
        const uint64_t bs01 = pk2(s_bn[ch0], s_bn[ch0 + 1]), bs23 = pk2(s_bn[ch0 + 2], s_bn[ch0 + 3]);
        // split output: the power-of-two operand scale rides on the BN constants (exact: relu(y) * sA == relu(y * sA))
        const uint64_t ba01 = pk2(s_bn[HC + ch0] * sA, s_bn[HC + ch0 + 1] * sA), ba23 = pk2(s_bn[HC + ch0 + 2] * sA, s_bn[HC + ch0 + 3] * sA);
        const uint64_t bb01 = pk2(s_bn[2 * HC + ch0] * sA, s_bn[2 * HC + ch0 + 1] * sA), bb23 = pk2(s_bn[2 * HC + ch0 + 2] * sA, s_bn[2 * HC + ch0 + 3] * sA);
        const float* base = rows + (pr * HW) * C + l16 * 4;
        const float* abase = a_sm + pr * G::AROW;
        ulonglong2 win[3][3];                                   // .x = channels (0, 1), .y = channels (2, 3)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            win[dy][1] = *reinterpret_cast<const ulonglong2*>(base + (dy * HW + 0) * C);
            win[dy][2] = *reinterpret_cast<const ulonglong2*>(base + (dy * HW + 1) * C);
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                win[dy][0] = win[dy][1]; win[dy][1] = win[dy][2];
                win[dy][2] = *reinterpret_cast<const ulonglong2*>(base + (dy * HW + c + 2) * C);
            }
            const ulonglong2 w01 = *reinterpret_cast<const ulonglong2*>(abase + c * AST);          // (w0,w0) (w1,w1)
            const ulonglong2 w23 = *reinterpret_cast<const ulonglong2*>(abase + c * AST + 4);
            const ulonglong2 w45 = *reinterpret_cast<const ulonglong2*>(abase + c * AST + 8);
            const ulonglong2 w67 = *reinterpret_cast<const ulonglong2*>(abase + c * AST + 12);
            const unsigned long long w8 = *reinterpret_cast<const unsigned long long*>(abase + c * AST + 16);
            const unsigned long long aw[9] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y, w67.x, w67.y, w8};
            unsigned long long acc01 = 0ull, acc23 = 0ull;
#pragma unroll
            for (int q = 0; q < O::N; ++q) {
                const int pz = O::pos(q);
                acc01 = ffma2(aw[pz], win[pz / 3][pz % 3].x, acc01);
                acc23 = ffma2(aw[pz], win[pz / 3][pz % 3].y, acc23);
            }
            const int id2 = ids[(pr + 1) * HW + c + 1];
            float y0, y1, y2, y3;
            upk2(ffma2(fadd2(acc01, bs01), ba01, bb01), y0, y1);
            upk2(ffma2(fadd2(acc23, bs23), ba23, bb23), y2, y3);
            if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
            if (SO.hi) {
                const __half2 h01 = __floats2half2_rn(y0, y1), h23 = __floats2half2_rn(y2, y3);     // already scaled by sA
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn((y0 - f01.x) * 2048.f, (y1 - f01.y) * 2048.f);
                const __half2 l23 = __floats2half2_rn((y2 - f23.x) * 2048.f, (y3 - f23.y) * 2048.f);
                if (id2 >= 0) {
                    uint16_t* prow = SO.hi + (int64_t)id2 * (2 * HC) + ch0;
                    *reinterpret_cast<uint2*>(prow) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                    *reinterpret_cast<uint2*>(prow + HC) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                }
            } else if (id2 >= 0) {
                *reinterpret_cast<float4*>(out + (int64_t)id2 * HC + ch0) = make_float4(y0, y1, y2, y3);
            }
        }
        } else {
        // 4. aggregate: warp w walks patch rows 2w, 2w+1; lane <-> VEC channels; 3x3 sliding window in registers
        float bs[VEC], ba[VEC], bb[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int ch = hh * C + lane * VEC + v;
            // split output: the power-of-two operand scale rides on the BN constants (exact: relu(y) * sA == relu(y * sA))
            bs[v] = s_bn[ch]; ba[v] = s_bn[HC + ch] * sA; bb[v] = s_bn[2 * HC + ch] * sA;
        }
        auto ldrow = [&](const float* p, float* dst) {
            if (VEC == 1) dst[0] = p[0];
            else if (VEC == 2) { const float2 t = *reinterpret_cast<const float2*>(p); dst[0] = t.x; dst[1 % VEC] = t.y; }
            else { const float4 t = *reinterpret_cast<const float4*>(p); dst[0] = t.x; dst[1 % VEC] = t.y; dst[2 % VEC] = t.z; dst[3 % VEC] = t.w; }
        };
#pragma unroll 1
        for (int rr = 0; rr < 2; ++rr) {
            const int pr = warp * 2 + rr;                                   // patch row; halo rows pr .. pr+2
            const float* base = rows + (pr * HW) * C + lane * VEC;
            float win[3][3][VEC];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                ldrow(base + (dy * HW + 0) * C, win[dy][1]);
                ldrow(base + (dy * HW + 1) * C, win[dy][2]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { win[dy][0][v] = win[dy][1][v]; win[dy][1][v] = win[dy][2][v]; }
                    ldrow(base + (dy * HW + c + 2) * C, win[dy][2]);
                }
                const float4 a0 = *reinterpret_cast<const float4*>(a_sm + pr * G::AROW + c * AST);
                const float4 a1 = *reinterpret_cast<const float4*>(a_sm + pr * G::AROW + c * AST + 4);
                const float aw[9] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a_sm[pr * G::AROW + c * AST + 8]};
                float acc[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll
                for (int q = 0; q < O::N; ++q) {
                    const int p = O::pos(q);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[v] = fmaf(aw[p], win[p / 3][p % 3][v], acc[v]);
                }
                const int id2 = ids[(pr + 1) * HW + c + 1];
                if (id2 >= 0) {
                    float o[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        float y = acc[v] + bs[v];
                        y = fmaf(y, ba[v], bb[v]);
                        o[v] = relu ? fmaxf(y, 0.f) : y;
                    }
                    if (VEC % 2 == 0 && SO.hi) {
                        uint16_t* prow = SO.hi + (int64_t)id2 * (2 * HC) + hh * C + lane * VEC;
                        const int lo_off = HC;
#pragma unroll
                        for (int v = 0; v + 1 < VEC; v += 2) {
                            const float s0v = o[v], s1v = o[v + 1];             // already scaled by sA
                            const __half2 hi2 = __floats2half2_rn(s0v, s1v);
                            const float2 hf = __half22float2(hi2);
                            const __half2 lo2 = __floats2half2_rn((s0v - hf.x) * 2048.f, (s1v - hf.y) * 2048.f);
                            *reinterpret_cast<__half2*>(prow + v) = hi2;
                            *reinterpret_cast<__half2*>(prow + lo_off + v) = lo2;
                        }
                    } else {
                        VecIO<VEC>::st(out + (int64_t)id2 * HC + hh * C + lane * VEC, o);
                    }
                }
            }
        }
        }
        AGG_TICK(4);
        __syncthreads();                     // every warp is done with rows / a_sm before the next head overwrites them
        AGG_TICK(5);
    }
}

template <int H, int C, int PH, bool CONN4>
int launch_agg_tile_t(const LayerDev& L, const float* h, const float* alpha, const TopoDev& T, const float* grid_eattr,
                      int ED, int relu, float slope, float* out, int64_t n_nodes, const SplitOut& SO, cudaStream_t s) {
    using G = TileCfg<H, C, PH>;
    static_assert(G::SMEM <= 232448, "shared memory budget exceeded");
    auto kern = gat_aggregate_tile<H, C, PH, CONN4>;
    BGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));   // per device, idempotent, thread-safe
    dim3 grid((unsigned)((T.cols + 15) / 16), (unsigned)((T.rows + PH - 1) / PH));
    kern<<<grid, PH * 16, G::SMEM, s>>>(h, alpha, T, grid_eattr, ED, L.Medge, L.bias, L.bn_a, L.bn_b, relu, slope, out, n_nodes, SO);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}


static bool tile_supported(const TopoDev& T, int H, int C, int ED, const float* grid_eattr, bool* conn4);

// ------------------------------------------------------------------------------------------ ragged band-patch aggregation
// Ragged batches (BASELINE configs[3]: thousands of small refinement grids packed end to end) with the TILE kernel's data
// flow: a patch is a band of rows of ONE grid over its full width (geometry per patch from a table the host plans:
// (rows + 2) x (width + 2) <= 324 halo slots, rows x width <= 256 destinations), so a halo row is again a run of cells with
// consecutive node ids = one bulk copy per head, absent cells are zero rows, and warp w walks patch rows w, w + 8, ... with the
// 3 x 3 sliding register window (3 new LDS.64 per destination).  The windowed CSR kernel this replaces staged a node-id WINDOW
// (1.8x read amplification at width 50) and gathered nine rows per destination from shared memory by source id: 6.1 ms per
// H = 4 layer of a 7.2 M-node chunk = 39 % of its byte roofline.  Same arithmetic and edge order as gat_aggregate_tile.
struct RagPatch { long long cell_off; int gh, gw, r0, nr; };       // grid cell offset, grid rows / cols, first row, rows of the band

template <int H, int C> struct RagCfg {
    static constexpr int NSLOT = 324, NT = 256, AST = 12, RUNS = 256;
    static constexpr size_t SMEM = (size_t)(NSLOT * C + NSLOT + NSLOT * H + NT * AST + BGNN_MAX_EDGE_FEATS * H + 3 * H * C + 2 * RUNS + 64) * 4 + 32;
    static constexpr bool ALPHA_BULK = (2 * H * 4) % 16 == 0 && NSLOT * 2 * H <= NT * AST;
};

template <int H, int C, bool CONN4>
__global__ void __launch_bounds__(256, 2)
gat_aggregate_ragged(const float* __restrict__ h, const float* __restrict__ alpha, const int32_t* __restrict__ node_id,
                     const RagPatch* __restrict__ patches, const TopoDev T, const float* __restrict__ grid_eattr, int ED,
                     const float* __restrict__ Medge, const float* __restrict__ bias, const float* __restrict__ bn_a,
                     const float* __restrict__ bn_b, int relu, float slope, float* __restrict__ out, int64_t n_nodes,
                     const SplitOut SO) {
    using G = RagCfg<H, C>;
    using O = EdgeOrder<CONN4>;
    constexpr int NSLOT = G::NSLOT, NT = G::NT, AST = G::AST;
    constexpr int VEC = C / 32, HC = H * C;
    static_assert(VEC == 2, "lane <-> 2 channels");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* rows = reinterpret_cast<float*>(smem_raw);          // [NSLOT][C]   source rows of the current head
    int* ids = reinterpret_cast<int*>(rows + NSLOT * C);       // [NSLOT]      node id per halo cell (-1 invalid)
    float* asrc = reinterpret_cast<float*>(ids + NSLOT);       // [H][NSLOT]   alpha_src per halo cell
    float* a_sm = asrc + NSLOT * H;                            // [NT][AST]    3x3 attention weights of the current head
    float* al_sm = a_sm;                                       // [NSLOT][2H]  landing area of the alpha bulk copies (setup only)
    float* M_s = a_sm + NT * AST;                              // [ED_MAX][H]
    float* s_bn = M_s + BGNN_MAX_EDGE_FEATS * H;               // [3][HC]
    int* run_tab = reinterpret_cast<int*>(s_bn + 3 * HC);      // [RUNS][2]    (start << 16 | length, node id), `cap` entries per halo row
    int* run_cnt = run_tab + 2 * G::RUNS;                      // [64]         runs per halo row
    uint64_t* bar = reinterpret_cast<uint64_t*>(run_cnt + 64); // [0] source rows, [1] alpha

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const RagPatch pt = patches[blockIdx.x];
    const int W = pt.gw, PH = pt.nr, HWr = W + 2, HRr = PH + 2, nslot = HRr * HWr, ndest = PH * W, cap = HWr / 2 + 1;
    const float sA = SO.hi ? operand_scale(SO.amax_src, SO.mul, SO.add) : 1.f;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(NT));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar + 1)), "r"(HRr));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int sl = tid; sl < nslot; sl += NT) {
        const int r = pt.r0 - 1 + sl / HWr, c = sl % HWr - 1;
        int id = -1;
        if (r >= 0 && r < pt.gh && c >= 0 && c < W) id = __ldg(node_id + pt.cell_off + (long long)r * W + c);
        ids[sl] = id;
    }
    for (int i = tid; i < HC; i += NT) { s_bn[i] = __ldg(bias + i); s_bn[HC + i] = __ldg(bn_a + i); s_bn[2 * HC + i] = __ldg(bn_b + i); }
    if (tid < BGNN_MAX_EDGE_FEATS * H) M_s[tid] = (tid / H) < ED ? __ldg(Medge + tid) : 0.f;
    const bool has_dest = tid < ndest;
    const int dr_ = has_dest ? tid / W : 0, dc_ = has_dest ? tid - dr_ * W : 0;
    const int cs = (dr_ + 1) * HWr + (dc_ + 1);
    float ea[8][BGNN_MAX_EDGE_FEATS];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ea[k][d] = 0.f;
    if (has_dest && ED > 0) {
        const float* ge = grid_eattr + (pt.cell_off + (long long)(pt.r0 + dr_) * W + dc_) * (T.K * ED);
        if (T.K == 8 && ED == 3) {
            float v[24];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(ge) + q);
                v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int d = 0; d < 3; ++d) ea[k][d] = v[k * 3 + d];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d)
                    if (k < T.K && d < ED) ea[k][d] = __ldg(ge + k * ED + d);
        }
    }
    __syncthreads();

    if (tid < HRr) {                               // runs of valid cells of halo row tid
        const int s0 = tid * HWr;
        int run = -1, n = 0;
        for (int c = 0; c <= HWr; ++c) {
            const int id = c < HWr ? ids[s0 + c] : -1;
            if (id >= 0 && run < 0) run = c;
            if (id < 0 && run >= 0) {
                run_tab[2 * (tid * cap + n)] = (run << 16) | (c - run);
                run_tab[2 * (tid * cap + n) + 1] = ids[s0 + run];
                ++n;
                run = -1;
            }
        }
        run_cnt[tid] = n;
    }
    auto issue_rows = [&](int hh) {
        if (tid < HRr) {
            const float* hp = h + (int64_t)hh * n_nodes * C;
            const int s0 = tid * HWr, n = run_cnt[tid];
            for (int i = 0; i < n; ++i) {
                const int sl = run_tab[2 * (tid * cap + i)], id = run_tab[2 * (tid * cap + i) + 1];
                bulk_g2s(rows + (s0 + (sl >> 16)) * C, hp + (int64_t)id * C, (uint32_t)(sl & 0xffff) * C * 4, bar);
            }
        }
    };
    if (G::ALPHA_BULK && tid < HRr) {
        const int s0 = tid * HWr, n = run_cnt[tid];
        for (int i = 0; i < n; ++i) {
            const int sl = run_tab[2 * (tid * cap + i)], id = run_tab[2 * (tid * cap + i) + 1];
            bulk_g2s(al_sm + (s0 + (sl >> 16)) * 2 * H, alpha + (int64_t)id * 2 * H, (uint32_t)(sl & 0xffff) * 2 * H * 4, bar + 1);
        }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr(bar + 1)) : "memory");
    }
    issue_rows(0);

    const int myid = has_dest ? ids[cs] : -1;
    float lsum[BGNN_MAX_EDGE_FEATS], ad[H];
    unsigned nbv = 0;
#pragma unroll
    for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) lsum[d] = 0.f;
#pragma unroll
    for (int hh = 0; hh < H; ++hh) ad[hh] = 0.f;
    if (myid >= 0) {
        int deg = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < T.K && ids[cs - T.dr[k] * HWr - T.dc[k]] >= 0) {
                nbv |= 1u << k;
                ++deg;
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) lsum[d] += ea[k][d];
            }
#pragma unroll
        for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) lsum[d] = lsum[d] / (float)(deg > 0 ? deg : 1);
    }
    for (int sl = tid; sl < nslot; sl += NT)
        if (ids[sl] < 0) {
#pragma unroll
            for (int q = 0; q < C / 4; ++q) *reinterpret_cast<float4*>(rows + sl * C + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    if (G::ALPHA_BULK) {
        mbar_wait_parity(bar + 1, 0);
        for (int sl = tid; sl < nslot; sl += NT)
            if (ids[sl] >= 0) {
#pragma unroll
                for (int hh = 0; hh < H; ++hh) asrc[hh * NSLOT + sl] = al_sm[sl * 2 * H + hh];
            }
        if (myid >= 0) {
#pragma unroll
            for (int hh = 0; hh < H; ++hh) ad[hh] = al_sm[cs * 2 * H + H + hh];
        }
    } else {
        for (int sl = tid; sl < nslot; sl += NT) {
            const int id = ids[sl];
            if (id >= 0) {
#pragma unroll
                for (int hh = 0; hh < H; ++hh) asrc[hh * NSLOT + sl] = __ldg(alpha + (int64_t)id * (2 * H) + hh);
            }
        }
        if (myid >= 0) {
#pragma unroll
            for (int hh = 0; hh < H; ++hh) ad[hh] = __ldg(alpha + (int64_t)myid * (2 * H) + H + hh);
        }
    }
    __syncthreads();                 // asrc complete; the alpha landing area (aliases a_sm) is free
    int wpos[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) wpos[k] = (1 - T.dr[k]) * 3 + (1 - T.dc[k]);

#pragma unroll 1
    for (int hh = 0; hh < H; ++hh) {
        if (hh > 0) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            issue_rows(hh);
        }
        {
            float4* az = reinterpret_cast<float4*>(a_sm + tid * AST);
            az[0] = make_float4(0.f, 0.f, 0.f, 0.f); az[1] = make_float4(0.f, 0.f, 0.f, 0.f); az[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (myid >= 0) {
            float Mh[BGNN_MAX_EDGE_FEATS], adh = ad[0];
#pragma unroll
            for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) Mh[d] = M_s[d * H + hh];
#pragma unroll
            for (int q = 1; q < H; ++q) adh = (q == hh) ? ad[q] : adh;
            float sc[9], mx = -CUDART_INF_F;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                sc[k] = 0.f;
                if ((nbv >> k) & 1u) {
                    float ae = 0.f;
#pragma unroll
                    for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ae = fmaf(ea[k][d], Mh[d], ae);
                    float v = (asrc[hh * NSLOT + cs - T.dr[k] * HWr - T.dc[k]] + adh) + ae;
                    v = v > 0.f ? v : v * slope;
                    sc[k] = v;
                    mx = fmaxf(mx, v);
                }
            }
            {
                float ae = 0.f;
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ae = fmaf(lsum[d], Mh[d], ae);
                float v = (asrc[hh * NSLOT + cs] + adh) + ae;
                v = v > 0.f ? v : v * slope;
                sc[8] = v;
                mx = fmaxf(mx, v);
            }
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((nbv >> k) & 1u) { sc[k] = expf(sc[k] - mx); sum += sc[k]; }
            sc[8] = expf(sc[8] - mx);
            sum += sc[8];
            sum += 1e-16f;
            const float inv = 1.0f / sum;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((nbv >> k) & 1u) a_sm[tid * AST + wpos[k]] = sc[k] * inv;
            a_sm[tid * AST + 4] = sc[8] * inv;
        }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_addr(bar)) : "memory");
        mbar_wait_parity(bar, (uint32_t)(hh & 1));
        float bs[VEC], ba[VEC], bb[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int ch = hh * C + lane * VEC + v;
            bs[v] = s_bn[ch]; ba[v] = s_bn[HC + ch] * sA; bb[v] = s_bn[2 * HC + ch] * sA;
        }
        // work units = (patch row, segment of <= 16 columns): wide grids have few rows per band (4 at width 50), so whole rows per
        // warp would leave half the warps idle and the others with 50 serial steps; 16-column segments give every warp ~32 steps
        const int nseg = (W + 15) >> 4, units = PH * nseg;
#pragma unroll 1
        for (int u = warp; u < units; u += NT / 32) {
            const int pr = u / nseg, c0 = (u - pr * nseg) << 4;             // patch row (halo rows pr .. pr + 2), first column
            const float* base = rows + (pr * HWr + c0) * C + lane * VEC;
            float2 win[3][3];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                win[dy][1] = *reinterpret_cast<const float2*>(base + (dy * HWr + 0) * C);
                win[dy][2] = *reinterpret_cast<const float2*>(base + (dy * HWr + 1) * C);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int c = c0 + j;
                if (c >= W) break;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    win[dy][0] = win[dy][1]; win[dy][1] = win[dy][2];
                    win[dy][2] = *reinterpret_cast<const float2*>(base + (dy * HWr + j + 2) * C);
                }
                const int t2 = pr * W + c;
                const float4 a0 = *reinterpret_cast<const float4*>(a_sm + t2 * AST);
                const float4 a1 = *reinterpret_cast<const float4*>(a_sm + t2 * AST + 4);
                const float aw[9] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a_sm[t2 * AST + 8]};
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                for (int q = 0; q < O::N; ++q) {
                    const int pz = O::pos(q);
                    acc0 = fmaf(aw[pz], win[pz / 3][pz % 3].x, acc0);
                    acc1 = fmaf(aw[pz], win[pz / 3][pz % 3].y, acc1);
                }
                const int id2 = ids[(pr + 1) * HWr + c + 1];
                float y0 = fmaf(acc0 + bs[0], ba[0], bb[0]), y1 = fmaf(acc1 + bs[1], ba[1], bb[1]);
                if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                if (SO.hi) {
                    const __half2 hi2 = __floats2half2_rn(y0, y1);                 // already scaled by sA
                    const float2 hf = __half22float2(hi2);
                    const __half2 lo2 = __floats2half2_rn((y0 - hf.x) * 2048.f, (y1 - hf.y) * 2048.f);
                    if (id2 >= 0) {
                        uint16_t* prow = SO.hi + (int64_t)id2 * (2 * HC) + hh * C + lane * VEC;
                        *reinterpret_cast<__half2*>(prow) = hi2;
                        *reinterpret_cast<__half2*>(prow + HC) = lo2;
                    }
                } else if (id2 >= 0) {
                    *reinterpret_cast<float2*>(out + (int64_t)id2 * HC + hh * C + lane * VEC) = make_float2(y0, y1);
                }
            }
        }
        __syncthreads();                     // every warp is done with rows / a_sm before the next head overwrites them
    }
}

// returns 1 when the band-patch kernel does not apply
int launch_agg_ragged(const LayerDev& L, const float* h, const float* alpha, const int32_t* node_id, const void* patches,
                      int n_patches, const TopoDev& T, const float* grid_eattr, int ED, int relu, float slope, float* out,
                      int64_t n_nodes, const SplitOut& SO, cudaStream_t s) {
    bool is4 = false;
    TopoDev T1 = T; T1.rows = 1; T1.cols = 1;
    if (L.C != 64 || !(L.H == 1 || L.H == 2 || L.H == 4) || ED > BGNN_MAX_EDGE_FEATS || !grid_eattr || !patches || !node_id) return 1;
    if (!tile_supported(T1, L.H, L.C, ED, grid_eattr, &is4)) return 1;
    if (n_patches == 0) return BGNN_OK;
#define RLAUNCH(H_, C4_)                                                                                                \
    do {                                                                                                                \
        using G = RagCfg<H_, 64>;                                                                                       \
        static_assert(G::SMEM <= 115000, "two CTAs per SM");                                                            \
        auto kern = gat_aggregate_ragged<H_, 64, C4_>;                                                                  \
        BGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));               \
        kern<<<(unsigned)n_patches, 256, G::SMEM, s>>>(h, alpha, node_id, reinterpret_cast<const RagPatch*>(patches), T, grid_eattr,  \
                                                       ED, L.Medge, L.bias, L.bn_a, L.bn_b, relu, slope, out, n_nodes, SO); \
    } while (0)
    if (is4) { if (L.H == 4) RLAUNCH(4, true); else if (L.H == 2) RLAUNCH(2, true); else RLAUNCH(1, true); }
    else { if (L.H == 4) RLAUNCH(4, false); else if (L.H == 2) RLAUNCH(2, false); else RLAUNCH(1, false); }
#undef RLAUNCH
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

// ------------------------------------------------------------------------------------------ windowed CSR aggregation
// Ragged batches (scripts/inference_native.py:249-342: thousands of small refinement grids packed end to end) have no
// single grid to tile, but their edges are LOCAL in node-id space: nodes are numbered in raster order per grid, so every
// source of a destination lies within (grid width + 1) ids of it.  A CTA takes 128 CONSECUTIVE destinations, finds the id
// range [lo, hi] their in-edges (and self loops) touch, and -- because h is head-major ([H][N][C]) -- stages that range of
// one head as ONE contiguous bulk copy in shared memory; destinations then read their source rows from shared memory
// instead of gathering 9 x 1 KB per destination through L1 / L2 (gat_aggregate).  Thread t owns destination t for the
// scores + softmax of all heads (computed once; weights and source ids parked in shared memory), warp w then walks 16
// destinations per head with lane <-> channel pair, FMAs in CSR order, self loop last -- the arithmetic of the tiled grid
// kernel, including its optional fp16 hi | lo output.  A block whose range does not fit (wide grids, foreign graphs) reads
// its rows from global memory instead.  Requires in-degree <= 8 (the caller's stencil hint) and C = 64.
template <int H> struct WinCfg {
    static constexpr int D = 128, MAXD = 8, SLOTS = MAXD + 1, SP = 12, C = 64, NT = 256;     // SP: padded slot stride (16 B rows)
    static constexpr int RCAP = 320;                                       // rows of one head kept in shared memory (80 KB)
    static constexpr size_t SMEM = (size_t)RCAP * C * 4 + (size_t)D * H * SP * 4 + (size_t)D * SP * 4 + D * 4 + 64;
};

template <int H>
__global__ void __launch_bounds__(256, 2)
gat_aggregate_window(const float* __restrict__ h, const float* __restrict__ alpha, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ csr_src, const float* __restrict__ eattr, int ED,
                     const float* __restrict__ Medge, const float* __restrict__ bias, const float* __restrict__ bn_a,
                     const float* __restrict__ bn_b, int relu, float slope, float* __restrict__ out, int64_t N,
                     const SplitOut SO) {
    using G = WinCfg<H>;
    constexpr int D = G::D, SLOTS = G::SLOTS, SP = G::SP, C = G::C, HC = H * C, RCAP = G::RCAP;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* rows = reinterpret_cast<float*>(smem_raw);                 // [RCAP][C]       source rows lo .. hi of the current head
    float* w_sm = rows + RCAP * C;                                    // [D][H][SP]      attention weights (slot order = CSR order, self loop
                                                                      //                 last; unused slots: weight 0 on the node itself)
    int* j_sm = reinterpret_cast<int*>(w_sm + D * H * SP);            // [D][SP]         source ids
    int* cnt_sm = j_sm + D * SP;                                      // [D]             in-degree + 1 (0: no such destination)
    int* s_rng = cnt_sm + D;                                          // lo, hi
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_rng + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t d0 = (int64_t)blockIdx.x * D;
    const float sA = SO.hi ? operand_scale(SO.amax_src, SO.mul, SO.add) : 1.f;
    if (tid == 0) {
        s_rng[0] = 0x7fffffff; s_rng[1] = -1;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // ---- scores + softmax: threads t and t + 128 <-> destination d0 + t, each taking half of the heads
    {
        const int dl = tid & (D - 1), hsel = tid >> 7;
        constexpr int HPT = H >= 2 ? H / 2 : 1;                       // heads per thread
        const int64_t i = d0 + dl;
        int tot = 0;
        if (i < N) {
            const int start = __ldg(rowptr + i);
            const int deg = __ldg(rowptr + i + 1) - start;
            if (deg > G::MAXD) asm volatile("trap;\n");                // the caller's in-degree hint was wrong
            tot = deg + 1;
            int jmin = (int)i, jmax = (int)i;
            float ea[G::MAXD][BGNN_MAX_EDGE_FEATS], lsum[BGNN_MAX_EDGE_FEATS];
            int js[G::MAXD];
#pragma unroll
            for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) lsum[d] = 0.f;
#pragma unroll
            for (int e = 0; e < G::MAXD; ++e) {
                js[e] = (int)i;
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ea[e][d] = 0.f;
                if (e < deg) {
                    js[e] = __ldg(csr_src + start + e);
                    jmin = min(jmin, js[e]); jmax = max(jmax, js[e]);
#pragma unroll
                    for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d)
                        if (d < ED) { ea[e][d] = __ldg(eattr + (int64_t)(start + e) * ED + d); lsum[d] += ea[e][d]; }
                }
            }
#pragma unroll
            for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) lsum[d] = lsum[d] / (float)(deg > 0 ? deg : 1);
            if (hsel == 0) {
                atomicMin(&s_rng[0], jmin);
                atomicMax(&s_rng[1], jmax);
#pragma unroll
                for (int e = 0; e < SLOTS; ++e) j_sm[dl * SP + e] = (e < G::MAXD && e < deg) ? js[e < G::MAXD ? e : 0] : (int)i;
            }
#pragma unroll 1
            for (int hh = hsel * HPT; hh < (hsel + 1) * HPT && hh < H; ++hh) {
                float Mh[BGNN_MAX_EDGE_FEATS];
#pragma unroll
                for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) Mh[d] = d < ED ? __ldg(Medge + d * H + hh) : 0.f;
                const float adh = __ldg(alpha + i * (2 * H) + H + hh);
                float sc[SLOTS], mx = -CUDART_INF_F;
#pragma unroll
                for (int e = 0; e < G::MAXD; ++e) {
                    sc[e] = 0.f;
                    if (e < deg) {
                        float ae = 0.f;
#pragma unroll
                        for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ae = fmaf(ea[e][d], Mh[d], ae);
                        float v = (__ldg(alpha + (int64_t)js[e] * (2 * H) + hh) + adh) + ae;
                        v = v > 0.f ? v : v * slope;
                        sc[e] = v;
                        mx = fmaxf(mx, v);
                    }
                }
                {
                    float ae = 0.f;
#pragma unroll
                    for (int d = 0; d < BGNN_MAX_EDGE_FEATS; ++d) ae = fmaf(lsum[d], Mh[d], ae);
                    float v = (__ldg(alpha + i * (2 * H) + hh) + adh) + ae;
                    v = v > 0.f ? v : v * slope;
                    sc[G::MAXD] = v;
                    mx = fmaxf(mx, v);
                }
                float sum = 0.f;
#pragma unroll
                for (int e = 0; e < G::MAXD; ++e)
                    if (e < deg) { sc[e] = expf(sc[e] - mx); sum += sc[e]; }
                sc[G::MAXD] = expf(sc[G::MAXD] - mx);
                sum += sc[G::MAXD];
                sum += 1e-16f;
                const float inv = 1.0f / sum;
                float* wr = w_sm + (dl * H + hh) * SP;
#pragma unroll
                for (int e = 0; e < SLOTS; ++e) wr[e] = 0.f;
#pragma unroll
                for (int e = 0; e < G::MAXD; ++e)
                    if (e < deg) wr[e] = sc[e] * inv;
                wr[deg] = sc[G::MAXD] * inv;
            }
        }
        if (hsel == 0) cnt_sm[dl] = tot;
    }
    __syncthreads();
    const int lo = s_rng[0], hi = s_rng[1];
    const bool staged = hi >= lo && hi - lo + 1 <= RCAP;
#pragma unroll 1
    for (int hh = 0; hh < H; ++hh) {
        const float* hp = h + (int64_t)hh * N * C;                    // head-major plane
        if (staged && tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            const uint32_t bytes = (uint32_t)(hi - lo + 1) * C * 4;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(smem_addr(rows)), "l"(hp + (int64_t)lo * C), "r"(bytes), "r"(smem_addr(bar)) : "memory");
        }
        const int ch = hh * C + lane * 2;
        const float bs0 = __ldg(bias + ch), bs1 = __ldg(bias + ch + 1);
        const float ba0 = __ldg(bn_a + ch) * sA, ba1 = __ldg(bn_a + ch + 1) * sA;
        const float bb0 = __ldg(bn_b + ch) * sA, bb1 = __ldg(bn_b + ch + 1) * sA;
        if (staged) mbar_wait_parity(bar, (uint32_t)(hh & 1));
#pragma unroll 1
        for (int t = 0; t < D / 8; ++t) {
            const int dl = warp * (D / 8) + t;
            const int tot = cnt_sm[dl];
            if (tot == 0) continue;
            const int64_t i = d0 + dl;
            // all nine slots, unrolled: independent loads first, then the FMA chain in slot order (unused slots add w = 0
            // times the node's own finite row: a no-op)
            const int4 j0 = *reinterpret_cast<const int4*>(j_sm + dl * SP), j1 = *reinterpret_cast<const int4*>(j_sm + dl * SP + 4);
            const int jj[SLOTS] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w, j_sm[dl * SP + 8]};
            const float* wr = w_sm + (dl * H + hh) * SP;
            const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
            const float ww[SLOTS] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, wr[8]};
            float2 hv[SLOTS];
            if (staged) {
#pragma unroll
                for (int q = 0; q < SLOTS; ++q) hv[q] = *reinterpret_cast<const float2*>(rows + (jj[q] - lo) * C + lane * 2);
            } else {
#pragma unroll
                for (int q = 0; q < SLOTS; ++q) hv[q] = __ldg(reinterpret_cast<const float2*>(hp + (int64_t)jj[q] * C + lane * 2));
            }
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int q = 0; q < SLOTS; ++q) { a0 = fmaf(ww[q], hv[q].x, a0); a1 = fmaf(ww[q], hv[q].y, a1); }
            float o0 = fmaf(a0 + bs0, ba0, bb0), o1 = fmaf(a1 + bs1, ba1, bb1);
            if (relu) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
            if (SO.hi) {
                uint16_t* prow = SO.hi + i * (2 * HC) + hh * C + lane * 2;
                const __half2 hi2 = __floats2half2_rn(o0, o1);
                const float2 hf = __half22float2(hi2);
                const __half2 lo2 = __floats2half2_rn((o0 - hf.x) * 2048.f, (o1 - hf.y) * 2048.f);
                *reinterpret_cast<__half2*>(prow) = hi2;
                *reinterpret_cast<__half2*>(prow + HC) = lo2;
            } else {
                *reinterpret_cast<float2*>(out + i * HC + hh * C + lane * 2) = make_float2(o0, o1);
            }
        }
        __syncthreads();                     // every warp is done with `rows` before the next head's copy lands
    }
}

// returns 1 when the windowed kernel does not apply
int launch_agg_window(const LayerDev& L, const float* h, const float* alpha, const int32_t* rowptr, const int32_t* src,
                      const float* eattr, int ED, int relu, float slope, float* out, int64_t N, const SplitOut& SO,
                      cudaStream_t s) {
    if (L.C != 64 || !(L.H == 1 || L.H == 2 || L.H == 4) || ED > BGNN_MAX_EDGE_FEATS) return 1;
    if (N == 0) return BGNN_OK;
#define WLAUNCH(H_)                                                                                                     \
    do {                                                                                                                \
        using G = WinCfg<H_>;                                                                                           \
        static_assert(G::SMEM <= 115000, "two CTAs per SM");                                                            \
        auto kern = gat_aggregate_window<H_>;                                                                           \
        BGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));               \
        kern<<<(unsigned)ceil_div64(N, G::D), G::NT, G::SMEM, s>>>(h, alpha, rowptr, src, eattr, ED, L.Medge, L.bias,   \
                                                                   L.bn_a, L.bn_b, relu, slope, out, N, SO);            \
    } while (0)
    if (L.H == 4) WLAUNCH(4); else if (L.H == 2) WLAUNCH(2); else WLAUNCH(1);
#undef WLAUNCH
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

static bool tile_supported(const TopoDev& T, int H, int C, int ED, const float* grid_eattr, bool* conn4);

// the reference's own offset lists (graph_construction.py:79-87) in their order; a tiled instantiation exists for (H, C)
static bool tile_supported(const TopoDev& T, int H, int C, int ED, const float* grid_eattr, bool* conn4) {
    static const int r8[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, c8[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    static const int r4[4] = {-1, 1, 0, 0}, c4[4] = {0, 0, -1, 1};
    bool is8 = T.K == 8, is4 = T.K == 4;
    for (int k = 0; k < 8 && is8; ++k) is8 = T.dr[k] == r8[k] && T.dc[k] == c8[k];
    for (int k = 0; k < 4 && is4; ++k) is4 = T.dr[k] == r4[k] && T.dc[k] == c4[k];
    if (!is8 && !is4) return false;
    if (ED > 0 && !grid_eattr) return false;           // the dense per-cell edge features are required
    if (conn4) *conn4 = is4;
    const bool hok = H == 1 || H == 2 || H == 4 || H == 8;
    return (C == 32 || C == 64) ? hok : (C == 128 ? (H == 1 || H == 2 || H == 4) : false);
}

// returns 1 if no tiled instantiation fits (caller falls back to the CSR kernel)
int launch_agg_tile(const LayerDev& L, const float* h, const float* alpha, const TopoDev& T, const float* grid_eattr,
                    int ED, int relu, float slope, float* out, int64_t n_nodes, const SplitOut& SO, cudaStream_t s) {
    bool is4 = false;
    if (!tile_supported(T, L.H, L.C, ED, grid_eattr, &is4)) return 1;
    if (SO.hi && L.C % 64 != 0) return 1;              // the split store packs channel pairs
#define TCASE(H_, C_, PH_)                                                                                              \
    if (L.H == H_ && L.C == C_)                                                                                         \
        return is4 ? launch_agg_tile_t<H_, C_, PH_, true>(L, h, alpha, T, grid_eattr, ED, relu, slope, out, n_nodes, SO, s)   \
                   : launch_agg_tile_t<H_, C_, PH_, false>(L, h, alpha, T, grid_eattr, ED, relu, slope, out, n_nodes, SO, s)
#ifndef BGNN_AGG_PH_H4
#define BGNN_AGG_PH_H4 16              /* patch height of the H = 4, C = 64 instantiation (8: three CTAs of four warps per SM: 1.90-1.97 vs 1.90 ms on one box, not kept) */
#endif
    TCASE(4, 64, BGNN_AGG_PH_H4); TCASE(1, 64, 8); TCASE(2, 64, 16); TCASE(8, 64, 16);      // (H = 1: 8 x 16 patches, three CTAs per SM: 0.62 -> 0.60 ms)
    TCASE(1, 32, 16); TCASE(2, 32, 16); TCASE(4, 32, 16); TCASE(8, 32, 16);
    TCASE(1, 128, 8); TCASE(2, 128, 8); TCASE(4, 128, 8);
#undef TCASE
    return 1;
}

// ------------------------------------------------------------------------------------------ radius-2 tiled aggregation
// Grid path for stencils up to 5 x 5 (BASELINE configs[4]: k = 24) and for the bf16 mode on any grid stencil.  One CTA per
// 8 x 16 patch of cells; the h rows of its 12 x 20 halo are staged ONCE per head in shared memory as 256-byte slabs
// (128 bf16 or 64 fp32 channels per cell; cp.async, zero-filled for absent cells) instead of being gathered 25 times per
// destination through L1/L2 (the warp-per-destination kernel moved 25 KB per destination: 31 ms per layer at 2048^2).
// Two threads per destination compute the scores + softmax of the current head (edge features and the self-loop mean stay
// in registers across the head loop) while the slab is in flight and leave the attention weights in shared memory indexed
// by OFFSET (absent neighbours: weight 0); warp w then walks patch row w with lane <-> channels, FMAs in the reference's
// edge order (offset k ascending, source = i - off_k, self loop last).  No per-edge tensor is materialised.
struct Topo5Dev {
    int rows, cols;
    const int32_t* node_id;
    int K;
    int8_t dr[BGNN_MAX_OFFSETS], dc[BGNN_MAX_OFFSETS];
    int16_t doff[BGNN_MAX_OFFSETS];        // halo slot of the source relative to the destination's: -dr * HW - dc
    bool window;                           // full 5x5 stencil: sliding register window variant allowed (ModelOpts::tile5_window)
};

template <int C, int IO, bool FULL5> struct Tile5Cfg {
    static constexpr int PH = 8, PW = 16, R = 2, HR = PH + 2 * R, HW = PW + 2 * R, NSLOT = HR * HW, ND = PH * PW, NT = 256;
    static constexpr int ESZ = IO == IO_F32 ? 4 : 2;                       // bytes per stored h element
    // channels per slab: 256-byte rows; the sliding-window variant keeps a 5 x 5 register window per lane and therefore
    // works on 64-channel slabs (2 channels per lane: 50 window registers, two or three CTAs per SM)
    static constexpr int CPS = FULL5 ? 64 : (C * ESZ <= 256 ? C : 256 / ESZ);
    static constexpr int RB = CPS * ESZ;                                   // slab row bytes (128 or 256)
    static constexpr int NSLAB = C / CPS;
    static constexpr int VEC = CPS / 32;                                   // channels per lane
    static constexpr int WST = BGNN_MAX_OFFSETS + 4;                       // weights per destination (24 offsets + self, 16 B aligned rows)
    static constexpr size_t SMEM = (size_t)NSLOT * RB + (size_t)ND * WST * 4 + (size_t)NSLOT * 4 + 16;
};

template <int H, int C, int IO, bool FULL5>
__global__ void __launch_bounds__(256, 2)
gat_aggregate_tile5(const float* __restrict__ h_, const float* __restrict__ alpha, const int32_t* __restrict__ rowptr,
                    const float* __restrict__ eattr, int ED, const float* __restrict__ Medge, const float* __restrict__ bias,
                    const float* __restrict__ bn_a, const float* __restrict__ bn_b, int relu, float slope,
                    float* __restrict__ out_, int64_t N, const __grid_constant__ Topo5Dev T) {
    using G = Tile5Cfg<C, IO, FULL5>;
    constexpr int HC = H * C, KH = (BGNN_MAX_OFFSETS + 1) / 2;            // offsets per thread of a pair (<= 12) + self
    extern __shared__ __align__(16) uint8_t smem5[];
    uint8_t* slab = smem5;                                                 // [NSLOT][RB]
    float* wts = reinterpret_cast<float*>(slab + (size_t)G::NSLOT * G::RB);   // [ND][WST]
    int32_t* hid = reinterpret_cast<int32_t*>(wts + G::ND * G::WST);      // [NSLOT] node id of the halo cell (-1 absent)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.y * G::PH, c0 = blockIdx.x * G::PW;
    const uint8_t* hb = reinterpret_cast<const uint8_t*>(h_);

    for (int p = tid; p < G::NSLOT; p += G::NT) {
        const int r = r0 - G::R + p / G::HW, c = c0 - G::R + p % G::HW;
        hid[p] = (r >= 0 && r < T.rows && c >= 0 && c < T.cols) ? __ldg(T.node_id + (int64_t)r * T.cols + c) : -1;
    }
    __syncthreads();

    auto stage = [&](int head, int sl) {                                   // slab (head, sl) of every halo cell -> shared memory
        constexpr int CPR = G::RB / 16;                                    // 16-byte pieces per row
        for (int q = tid; q < G::NSLOT * CPR; q += G::NT) {
            const int p = q / CPR, piece = q % CPR;
            const int id = hid[p];
            const uint8_t* src = id >= 0 ? hb + (((int64_t)head * N + id) * C + sl * G::CPS) * G::ESZ + piece * 16 : hb;
            const uint32_t dst = smem_addr(slab + (size_t)p * G::RB + piece * 16);
            const int nbytes = id >= 0 ? 16 : 0;                           // src-size 0: the 16 bytes are zero-filled
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    // ---- score set-up: thread pair (d, half); half handles offsets k = half, half + 2, ...; half 0 also the self loop
    const int d = tid >> 1, half = tid & 1;
    const int pr = d / G::PW, pc = d % G::PW;
    const int my_id = hid[(pr + G::R) * G::HW + pc + G::R];
    uint32_t vmask = 0;
    int start = 0;
    float ea[KH][BGNN_MAX_EDGE_FEATS];
    int jsrc[KH];
    float mean[BGNN_MAX_EDGE_FEATS];
#pragma unroll
    for (int f = 0; f < BGNN_MAX_EDGE_FEATS; ++f) mean[f] = 0.f;
    if (my_id >= 0) {
        for (int k = 0; k < T.K; ++k) {
            const int hp = (pr + G::R - T.dr[k]) * G::HW + pc + G::R - T.dc[k];
            if (hid[hp] >= 0) vmask |= 1u << k;
        }
        start = __ldg(rowptr + my_id);
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) {
            const int k = 2 * kk + half;
            jsrc[kk] = -1;
#pragma unroll
            for (int f = 0; f < BGNN_MAX_EDGE_FEATS; ++f) ea[kk][f] = 0.f;
            if (k < T.K && (vmask >> k & 1u)) {
                jsrc[kk] = hid[(pr + G::R - T.dr[k]) * G::HW + pc + G::R - T.dc[k]];
                const int e = start + __popc(vmask & ((1u << k) - 1u));
#pragma unroll
                for (int f = 0; f < BGNN_MAX_EDGE_FEATS; ++f)
                    if (f < ED) { ea[kk][f] = __ldg(eattr + (int64_t)e * ED + f); mean[f] += ea[kk][f]; }
            }
        }
    } else {
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) jsrc[kk] = -1;
    }
    {   // self-loop attribute = mean of the incoming edge attributes (0 when the node has none)
        const int deg = __popc(vmask);
#pragma unroll
        for (int f = 0; f < BGNN_MAX_EDGE_FEATS; ++f) {
            mean[f] += __shfl_xor_sync(0xffffffffu, mean[f], 1);
            mean[f] = mean[f] / (float)(deg > 0 ? deg : 1);
        }
    }

    for (int head = 0; head < H; ++head) {
        stage(head, 0);
        // ---- scores + softmax of this head -> wts[d][k] (k = offset index, 24 = self)
        {
            float M[BGNN_MAX_EDGE_FEATS];
#pragma unroll
            for (int f = 0; f < BGNN_MAX_EDGE_FEATS; ++f) M[f] = f < ED ? __ldg(Medge + f * H + head) : 0.f;
            float sc[KH], sself = -CUDART_INF_F, mx = -CUDART_INF_F;
            const float ad = my_id >= 0 ? __ldg(alpha + (int64_t)my_id * (2 * H) + H + head) : 0.f;
#pragma unroll
            for (int kk = 0; kk < KH; ++kk) {
                sc[kk] = -CUDART_INF_F;
                if (jsrc[kk] >= 0) {
                    float ae = 0.f;
#pragma unroll
                    for (int f = 0; f < BGNN_MAX_EDGE_FEATS; ++f) ae = fmaf(ea[kk][f], M[f], ae);
                    const float v = (__ldg(alpha + (int64_t)jsrc[kk] * (2 * H) + head) + ad) + ae;
                    sc[kk] = v > 0.f ? v : v * slope;
                    mx = fmaxf(mx, sc[kk]);
                }
            }
            if (half == 0 && my_id >= 0) {
                float ae = 0.f;
#pragma unroll
                for (int f = 0; f < BGNN_MAX_EDGE_FEATS; ++f) ae = fmaf(mean[f], M[f], ae);
                const float v = (__ldg(alpha + (int64_t)my_id * (2 * H) + head) + ad) + ae;
                sself = v > 0.f ? v : v * slope;
                mx = fmaxf(mx, sself);
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            float sum = 0.f;
#pragma unroll
            for (int kk = 0; kk < KH; ++kk) { sc[kk] = jsrc[kk] >= 0 ? expf(sc[kk] - mx) : 0.f; sum += sc[kk]; }
            if (half == 0 && my_id >= 0) { sself = expf(sself - mx); sum += sself; } else sself = 0.f;
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += 1e-16f;
            float* w = wts + d * G::WST;
#pragma unroll
            for (int kk = 0; kk < KH; ++kk) {
                const int k = 2 * kk + half;
                if (k < BGNN_MAX_OFFSETS) w[k] = sc[kk] / sum;
            }
            if (half == 0) w[BGNN_MAX_OFFSETS] = sself / sum;
        }
        for (int sl = 0; sl < G::NSLAB; ++sl) {
            if (sl > 0) stage(head, sl);
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();                                               // slab landed, weights visible
            // ---- aggregate: warp <-> patch row, lane <-> VEC channels of the slab
            if constexpr (FULL5) {
                // full 5 x 5 stencil in raster order (OFFSETS_24): the warp slides a 5 x 5 register window along its patch
                // row -- 5 new source rows per destination instead of 25 -- and multiplies in the reference's edge order
                // (k ascending <-> window row 4..0, window column 4..0, centre = self loop last)
                static_assert(G::VEC == 2, "window variant works on 64-channel slabs");
                float win[5][5][2];
                float cb[2], ca[2], cc[2];
                const int ch0 = head * C + sl * G::CPS + lane * 2;
#pragma unroll
                for (int v = 0; v < 2; ++v) { cb[v] = __ldg(bias + ch0 + v); ca[v] = __ldg(bn_a + ch0 + v); cc[v] = __ldg(bn_b + ch0 + v); }
                auto load_col = [&](int x, float (*dst)[2], int wy) {
                    const uint8_t* row = slab + (size_t)((warp + wy) * G::HW + x) * G::RB + lane * (2 * G::ESZ);
                    if constexpr (IO == IO_F32) { const float2 t = *reinterpret_cast<const float2*>(row); dst[0][0] = t.x; dst[0][1] = t.y; }
                    else { const uint32_t t = *reinterpret_cast<const uint32_t*>(row); dst[0][0] = __uint_as_float(t << 16); dst[0][1] = __uint_as_float(t & 0xffff0000u); }
                };
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int wy = 0; wy < 5; ++wy) load_col(x, &win[wy][x % 5], wy);
#pragma unroll
                for (int pcw = 0; pcw < G::PW; ++pcw) {
#pragma unroll
                    for (int wy = 0; wy < 5; ++wy) load_col(pcw + 4, &win[wy][(pcw + 4) % 5], wy);
                    const int id = hid[(warp + G::R) * G::HW + pcw + G::R];
                    if (id >= 0) {                                         // warp-uniform
                        const float* w = wts + (warp * G::PW + pcw) * G::WST;
                        float wk[G::WST];
#pragma unroll
                        for (int q = 0; q < G::WST / 4; ++q) {
                            const float4 t = reinterpret_cast<const float4*>(w)[q];
                            wk[4 * q] = t.x; wk[4 * q + 1] = t.y; wk[4 * q + 2] = t.z; wk[4 * q + 3] = t.w;
                        }
                        float a0 = 0.f, a1 = 0.f;
                        int k = 0;
#pragma unroll
                        for (int wy = 4; wy >= 0; --wy)
#pragma unroll
                            for (int wx = 4; wx >= 0; --wx) {
                                if (wy == 2 && wx == 2) continue;
                                a0 = fmaf(wk[k], win[wy][(pcw + wx) % 5][0], a0);
                                a1 = fmaf(wk[k], win[wy][(pcw + wx) % 5][1], a1);
                                ++k;
                            }
                        a0 = fmaf(wk[BGNN_MAX_OFFSETS], win[2][(pcw + 2) % 5][0], a0);
                        a1 = fmaf(wk[BGNN_MAX_OFFSETS], win[2][(pcw + 2) % 5][1], a1);
                        float o[2];
                        o[0] = fmaf(a0 + cb[0], ca[0], cc[0]);
                        o[1] = fmaf(a1 + cb[1], ca[1], cc[1]);
                        if (relu) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); }
                        if constexpr (IO == IO_BF16) VecBF<2>::st(reinterpret_cast<uint16_t*>(out_) + (int64_t)id * HC + ch0, o);
                        else VecIO<2>::st(out_ + (int64_t)id * HC + ch0, o);
                    }
                }
            } else
            for (int pcw = 0; pcw < G::PW; ++pcw) {
                const int dd = warp * G::PW + pcw;
                const int id = hid[(warp + G::R) * G::HW + pcw + G::R];
                if (id < 0) continue;                                      // warp-uniform
                const float* w = wts + dd * G::WST;
                float acc[G::VEC];
#pragma unroll
                for (int v = 0; v < G::VEC; ++v) acc[v] = 0.f;
                auto fma_row = [&](int hp, float a) {
                    const uint8_t* row = slab + (size_t)hp * G::RB + lane * (G::VEC * G::ESZ);
                    if constexpr (IO == IO_F32) {
                        if constexpr (G::VEC == 2) { const float2 t = *reinterpret_cast<const float2*>(row); acc[0] = fmaf(a, t.x, acc[0]); acc[1] = fmaf(a, t.y, acc[1]); }
                        else { acc[0] = fmaf(a, *reinterpret_cast<const float*>(row), acc[0]); }
                    } else {
                        if constexpr (G::VEC == 4) {
                            const uint2 t = *reinterpret_cast<const uint2*>(row);
                            acc[0] = fmaf(a, __uint_as_float(t.x << 16), acc[0]); acc[1] = fmaf(a, __uint_as_float(t.x & 0xffff0000u), acc[1]);
                            acc[2] = fmaf(a, __uint_as_float(t.y << 16), acc[2]); acc[3] = fmaf(a, __uint_as_float(t.y & 0xffff0000u), acc[3]);
                        } else if constexpr (G::VEC == 2) {
                            const uint32_t t = *reinterpret_cast<const uint32_t*>(row);
                            acc[0] = fmaf(a, __uint_as_float(t << 16), acc[0]); acc[1] = fmaf(a, __uint_as_float(t & 0xffff0000u), acc[1]);
                        } else {
                            acc[0] = fmaf(a, bf16_bits_to_f32(*reinterpret_cast<const uint16_t*>(row)), acc[0]);
                        }
                    }
                };
                const int hp0 = (warp + G::R) * G::HW + pcw + G::R;
                float wk[G::WST];                                          // this destination's weights: 7 x LDS.128 (broadcast)
#pragma unroll
                for (int q = 0; q < G::WST / 4; ++q) {
                    const float4 t = reinterpret_cast<const float4*>(w)[q];
                    wk[4 * q] = t.x; wk[4 * q + 1] = t.y; wk[4 * q + 2] = t.z; wk[4 * q + 3] = t.w;
                }
#pragma unroll
                for (int k = 0; k < BGNN_MAX_OFFSETS; ++k)
                    if (k < T.K) fma_row(hp0 + T.doff[k], wk[k]);
                fma_row(hp0, wk[BGNN_MAX_OFFSETS]);
                // epilogue: + bias, BN(eval) as x*a + b, ReLU
                float o[G::VEC];
                const int ch0 = head * C + sl * G::CPS + lane * G::VEC;
#pragma unroll
                for (int v = 0; v < G::VEC; ++v) {
                    float y = acc[v] + __ldg(bias + ch0 + v);
                    y = fmaf(y, __ldg(bn_a + ch0 + v), __ldg(bn_b + ch0 + v));
                    o[v] = relu ? fmaxf(y, 0.f) : y;
                }
                if constexpr (IO == IO_BF16) VecBF<G::VEC>::st(reinterpret_cast<uint16_t*>(out_) + (int64_t)id * HC + ch0, o);
                else VecIO<G::VEC>::st(out_ + (int64_t)id * HC + ch0, o);
            }
            __syncthreads();                                               // slab and weights may be overwritten
        }
    }
}

static bool topo5_to_dev(const bgnn_grid_topo* topo, int64_t E, Topo5Dev& T) {
    if (!(topo && topo->node_id && topo->rows > 0 && topo->n_offsets >= 1 && topo->n_offsets <= BGNN_MAX_OFFSETS && E > 0)) return false;
    T.rows = topo->rows; T.cols = topo->cols; T.node_id = topo->node_id; T.K = topo->n_offsets;
    for (int k = 0; k < BGNN_MAX_OFFSETS; ++k) {
        const int dr = k < T.K ? topo->off_dr[k] : 0, dc = k < T.K ? topo->off_dc[k] : 0;
        if (abs(dr) > 2 || abs(dc) > 2) return false;
        T.dr[k] = (int8_t)dr; T.dc[k] = (int8_t)dc;
        T.doff[k] = (int16_t)(-dr * 20 - dc);                 // Tile5Cfg::HW == 20
    }
    return true;
}

template <int H, int C, int IO, bool FULL5>
int launch_agg_tile5_tt(const LayerDev& L, const float* h, const float* alpha, const int32_t* rowptr, const float* eattr, int ED,
                        int relu, float slope, float* out, int64_t N, const Topo5Dev& T, cudaStream_t s) {
    using G = Tile5Cfg<C, IO, FULL5>;
    static_assert(G::SMEM <= 113000, "two CTAs per SM");
    auto kern = gat_aggregate_tile5<H, C, IO, FULL5>;
    BGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
    dim3 grid((unsigned)((T.cols + G::PW - 1) / G::PW), (unsigned)((T.rows + G::PH - 1) / G::PH));
    kern<<<grid, G::NT, G::SMEM, s>>>(h, alpha, rowptr, eattr, ED, L.Medge, L.bias, L.bn_a, L.bn_b, relu, slope, out, N, T);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

template <int H, int C, int IO>
int launch_agg_tile5_t(const LayerDev& L, const float* h, const float* alpha, const int32_t* rowptr, const float* eattr, int ED,
                       int relu, float slope, float* out, int64_t N, const Topo5Dev& T, cudaStream_t s) {
    bool full5 = T.window && T.K == BGNN_MAX_OFFSETS;          // exactly OFFSETS_24: raster order over the 5 x 5 window
    for (int k = 0, q = 0; full5 && k < T.K; ++k, ++q) {
        if (q == 12) ++q;                                             // the centre is not an offset
        full5 = T.dr[k] == q / 5 - 2 && T.dc[k] == q % 5 - 2;
    }
    if (full5) return launch_agg_tile5_tt<H, C, IO, true>(L, h, alpha, rowptr, eattr, ED, relu, slope, out, N, T, s);
    return launch_agg_tile5_tt<H, C, IO, false>(L, h, alpha, rowptr, eattr, ED, relu, slope, out, N, T, s);
}

// returns 1 when the radius-2 tiled kernel does not apply
int launch_agg_tile5(const LayerDev& L, const float* h, const float* alpha, const int32_t* rowptr, const float* eattr, int ED,
                     int relu, float slope, float* out, int64_t N, const Topo5Dev& T, int io, cudaStream_t s) {
    if (ED > BGNN_MAX_EDGE_FEATS || (ED > 0 && !eattr) || !rowptr) return 1;
    if (T.rows > 65535 * 8) return 1;
#define T5(H_, C_)                                                                                                              \
    if (L.H == H_ && L.C == C_) {                                                                                               \
        if (io == IO_F32) return launch_agg_tile5_t<H_, C_, IO_F32>(L, h, alpha, rowptr, eattr, ED, relu, slope, out, N, T, s); \
        if (io == IO_BF16) return launch_agg_tile5_t<H_, C_, IO_BF16>(L, h, alpha, rowptr, eattr, ED, relu, slope, out, N, T, s); \
        return launch_agg_tile5_t<H_, C_, IO_BF16_F32>(L, h, alpha, rowptr, eattr, ED, relu, slope, out, N, T, s);              \
    }
    T5(1, 64) T5(2, 64) T5(4, 64) T5(1, 128) T5(2, 128) T5(4, 128)
#undef T5
    return 1;
}

// ------------------------------------------------------------------------------------------ heads
// hh: [N, nblk*h2] = ReLU(first layers).  One warp per node.
__global__ void __launch_bounds__(256)
heads_finish(const float* __restrict__ hh, int h2, int nblk, int ncls, const float* __restrict__ Wcls,
             const float* __restrict__ bcls, const float* __restrict__ Wconf, const float* __restrict__ bconf,
             const float* __restrict__ Wcorr, const float* __restrict__ bcorr, float* __restrict__ logits,
             float* __restrict__ probs, int64_t* __restrict__ cls, float* __restrict__ conf, float* __restrict__ corr,
             int64_t* __restrict__ action, float thr_auto, float thr_review, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = wid; n < N; n += nw) {
        const float* row = hh + n * (int64_t)(nblk * h2);
        float lg[8];
        float best = -CUDART_INF_F;
        for (int k = 0; k < ncls; ++k) {
            float p = 0.f;
            for (int u = lane; u < h2; u += 32) p = fmaf(__ldg(row + u), __ldg(Wcls + k * h2 + u), p);
            for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
            lg[k] = p + __ldg(bcls + k);
            best = fmaxf(best, lg[k]);
        }
        float pc = 0.f, pr = 0.f;
        for (int u = lane; u < h2; u += 32) {
            pc = fmaf(__ldg(row + h2 + u), __ldg(Wconf + u), pc);
            if (nblk > 2) pr = fmaf(__ldg(row + 2 * h2 + u), __ldg(Wcorr + u), pr);
        }
        for (int o = 16; o; o >>= 1) {
            pc += __shfl_xor_sync(0xffffffffu, pc, o);
            pr += __shfl_xor_sync(0xffffffffu, pr, o);
        }
        if (lane == 0) {
            float pb[8], sum = 0.f;
            for (int k = 0; k < ncls; ++k) { pb[k] = expf(lg[k] - best); sum += pb[k]; }
            int arg = 0; float pmax = -1.f;
            for (int k = 0; k < ncls; ++k) {
                pb[k] = pb[k] / sum;
                if (pb[k] > pmax) { pmax = pb[k]; arg = k; }      // first max wins (torch.argmax)
                if (logits) logits[n * ncls + k] = lg[k];
                if (probs) probs[n * ncls + k] = pb[k];
            }
            const float cf = 1.f / (1.f + expf(-(pc + __ldg(bconf))));
            if (cls) cls[n] = arg;
            if (conf) conf[n] = cf;
            if (corr && nblk > 2) corr[n] = pr + __ldg(bcorr);
            if (action) {
                int64_t a = (arg == 2 && cf > thr_auto) ? 1 : 0;          // gnn.py:439-441
                if (cf < thr_review) a = 2;                                // gnn.py:444-445
                action[n] = a;
            }
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------ model
static void layer_shape(const bgnn_model_desc* d, int l, int& in, int& H, int& C, int& HC, int& out) {
    const bool last = l == d->layers - 1;
    C = d->hidden;
    H = last ? 1 : d->heads;
    HC = H * C;
    in = l == 0 ? d->hidden : d->hidden * d->heads;
    out = last ? C : HC;
}

extern "C" int64_t bgnn_model_param_floats(const bgnn_model_desc* d) {
    if (!d) return -1;
    const int64_t C = d->hidden, F = d->in_channels, ED = d->edge_dim, h2 = C / 2;
    int64_t n = C * F + C + C * C + C;
    for (int l = 0; l < d->layers; ++l) {
        int in, H, Cc, HC, out;
        layer_shape(d, l, in, H, Cc, HC, out);
        n += (int64_t)HC * in + 2 * HC + (ED > 0 ? (int64_t)HC * ED + HC : 0) + 5 * (int64_t)out;
    }
    n += h2 * C + h2 + (int64_t)d->num_classes * h2 + d->num_classes;
    n += h2 * C + h2 + h2 + 1;
    if (d->predict_correction) n += h2 * C + h2 + h2 + 1;
    return n;
}

extern "C" int bgnn_model_create(const bgnn_model_desc* d, const float* p, int64_t n_floats, int device,
                                 bgnn_model** out) {
    BGNN_CHECK_ARG(d && p && out, "null argument");
    BGNN_CHECK_ARG(d->layers >= 1 && d->layers <= 16, "layers out of range");
    BGNN_CHECK_ARG(d->hidden == 32 || d->hidden == 64 || d->hidden == 128, "hidden must be 32, 64 or 128");
    BGNN_CHECK_ARG(d->heads == 1 || d->heads == 2 || d->heads == 4 || d->heads == 8, "heads must be 1, 2, 4 or 8");
    BGNN_CHECK_ARG(d->heads * d->hidden <= 512, "heads*hidden > 512 is not compiled");
    BGNN_CHECK_ARG(d->edge_dim >= 0 && d->edge_dim <= BGNN_MAX_EDGE_FEATS, "edge_dim out of range");
    BGNN_CHECK_ARG(d->num_classes >= 1 && d->num_classes <= 8, "num_classes out of range");
    BGNN_CHECK_ARG(d->in_channels >= 1 && d->in_channels <= 64, "in_channels out of range");
    BGNN_CHECK_ARG(n_floats == bgnn_model_param_floats(d), "parameter blob has %lld floats, expected %lld",
                   (long long)n_floats, (long long)bgnn_model_param_floats(d));
    const int C = d->hidden, F = d->in_channels, ED = d->edge_dim, h2 = C / 2;
    const int nblk = d->predict_correction ? 3 : 2;

    // ---- fold on the host, then one upload
    std::vector<float> hb;
    auto push = [&](const float* src, int64_t n) { size_t o = hb.size(); hb.insert(hb.end(), src, src + n); hb.resize(align_up(hb.size(), 4)); return o; };
    struct LOff { size_t W, as, ad, M, bias, a, b; };
    std::vector<LOff> lo(d->layers);
    const float* q = p;
    size_t oW0 = push(q, (int64_t)C * F); q += (int64_t)C * F;
    size_t ob0 = push(q, C); q += C;
    size_t oW1 = push(q, (int64_t)C * C); q += (int64_t)C * C;
    size_t ob1 = push(q, C); q += C;
    for (int l = 0; l < d->layers; ++l) {
        int in, H, Cc, HC, outc;
        layer_shape(d, l, in, H, Cc, HC, outc);
        lo[l].W = push(q, (int64_t)HC * in); q += (int64_t)HC * in;
        lo[l].as = push(q, HC); q += HC;
        lo[l].ad = push(q, HC); q += HC;
        std::vector<float> M((size_t)BGNN_MAX_EDGE_FEATS * H, 0.f);
        if (ED > 0) {
            const float* We = q; q += (int64_t)HC * ED;        // [HC, ED]
            const float* ae = q; q += HC;
            for (int e = 0; e < ED; ++e)
                for (int hh = 0; hh < H; ++hh) {
                    double s = 0.0;
                    for (int c = 0; c < Cc; ++c) s += (double)We[(int64_t)(hh * Cc + c) * ED + e] * (double)ae[hh * Cc + c];
                    M[(size_t)e * H + hh] = (float)s;
                }
        }
        lo[l].M = push(M.data(), (int64_t)M.size());
        lo[l].bias = push(q, outc); q += outc;
        const float *bw = q, *bb = q + outc, *bm = q + 2 * outc, *bv = q + 3 * outc;
        q += 4 * (int64_t)outc;
        std::vector<float> a(outc), b(outc);
        for (int c = 0; c < outc; ++c) {            // torch CPU eval BN: alpha = invstd*w, beta = b - mean*alpha
            const float invstd = 1.0f / sqrtf(bv[c] + d->bn_eps);
            a[c] = invstd * bw[c];
            b[c] = bb[c] - bm[c] * a[c];
        }
        lo[l].a = push(a.data(), outc);
        lo[l].b = push(b.data(), outc);
    }
    // heads: first layers concatenated [nblk*h2, C]
    const float* cW0 = q; q += (int64_t)h2 * C; const float* cb0 = q; q += h2;
    const float* cW1 = q; q += (int64_t)d->num_classes * h2; const float* cb1 = q; q += d->num_classes;
    const float* fW0 = q; q += (int64_t)h2 * C; const float* fb0 = q; q += h2;
    const float* fW1 = q; q += h2; const float* fb1 = q; q += 1;
    const float *rW0 = nullptr, *rb0 = nullptr, *rW1 = nullptr, *rb1 = nullptr;
    if (d->predict_correction) { rW0 = q; q += (int64_t)h2 * C; rb0 = q; q += h2; rW1 = q; q += h2; rb1 = q; q += 1; }
    std::vector<float> hW0((size_t)nblk * h2 * C), hb0((size_t)nblk * h2);
    memcpy(hW0.data(), cW0, sizeof(float) * h2 * C); memcpy(hb0.data(), cb0, sizeof(float) * h2);
    memcpy(hW0.data() + (size_t)h2 * C, fW0, sizeof(float) * h2 * C); memcpy(hb0.data() + h2, fb0, sizeof(float) * h2);
    if (rW0) { memcpy(hW0.data() + (size_t)2 * h2 * C, rW0, sizeof(float) * h2 * C); memcpy(hb0.data() + 2 * h2, rb0, sizeof(float) * h2); }
    size_t ohW0 = push(hW0.data(), (int64_t)hW0.size()), ohb0 = push(hb0.data(), (int64_t)hb0.size());
    size_t oWc = push(cW1, (int64_t)d->num_classes * h2), obc = push(cb1, d->num_classes);
    size_t oWf = push(fW1, h2), obf = push(fb1, 1);
    size_t oWr = rW1 ? push(rW1, h2) : 0, obr = rb1 ? push(rb1, 1) : 0;
    if (q - p != n_floats) { bgnn_set_error("internal: parameter walk mismatch"); return BGNN_E_INVALID; }

    // fp16 hi/lo splits of the projection weights for the tensor-core path
    std::vector<uint16_t> h16;
    std::vector<size_t> o16(d->layers, (size_t)-1);
    std::vector<float> winv(d->layers, 1.f);
    std::vector<std::array<size_t, 4>> o16blk(d->layers, std::array<size_t, 4>{(size_t)-1, (size_t)-1, (size_t)-1, (size_t)-1});
    std::vector<std::array<float, 4>> winvblk(d->layers, std::array<float, 4>{1.f, 1.f, 1.f, 1.f});
    for (int l = 0; l < d->layers; ++l) {
        int in, H, Cc, HC, outc;
        layer_shape(d, l, in, H, Cc, HC, outc);
        if (!proj_umma_supported(HC, in, Cc, true)) {
            if (HC == 512 && 128 % Cc == 0 && proj_umma_supported(128, in, Cc, true)) {      // column blocks of 128
                for (int b = 0; b < 4; ++b) {
                    o16blk[l][b] = h16.size();
                    h16.resize(h16.size() + (size_t)2 * 128 * in);
                    winvblk[l][b] = proj_umma_split_weights(hb.data() + lo[l].W + (size_t)b * 128 * in, 128, in, h16.data() + o16blk[l][b]);
                }
            }
            continue;
        }
        o16[l] = h16.size();
        h16.resize(h16.size() + (size_t)2 * HC * in);
        winv[l] = proj_umma_split_weights(hb.data() + lo[l].W, HC, in, h16.data() + o16[l]);
    }

    // layer 0 with LocalFeatureExtractor.mlp[3] folded in: h = (W1 r + b1) Wl^T = r (Wl W1)^T + Wl b1   (fp64 fold)
    size_t oFold16 = (size_t)-1, oFoldBias = 0;
    float fold0_inv = 1.f, heads_inv = 1.f;
    {
        int in, H, Cc, HC, outc;
        layer_shape(d, 0, in, H, Cc, HC, outc);
        if (proj_umma_supported(HC, in, Cc, true) && proj_umma_mlp0_supported(HC, C, F)) {
            std::vector<float> Wf((size_t)HC * C), bf(HC);
            const float* Wl = hb.data() + lo[0].W;      // [HC, C]
            const float* W1 = hb.data() + oW1;          // [C, C]
            const float* b1 = hb.data() + ob1;
            for (int n = 0; n < HC; ++n) {
                for (int k = 0; k < C; ++k) {
                    double acc = 0.0;
                    for (int c = 0; c < C; ++c) acc += (double)Wl[(size_t)n * C + c] * (double)W1[(size_t)c * C + k];
                    Wf[(size_t)n * C + k] = (float)acc;
                }
                double acc = 0.0;
                for (int c = 0; c < C; ++c) acc += (double)Wl[(size_t)n * C + c] * (double)b1[c];
                bf[n] = (float)acc;
            }
            oFoldBias = push(bf.data(), HC);
            oFold16 = h16.size();
            h16.resize(h16.size() + (size_t)2 * HC * C);
            fold0_inv = proj_umma_split_weights(Wf.data(), HC, C, h16.data() + oFold16);
        }
    }
    // unfused feature extractor (shapes without the PM_MLP0 fusion, e.g. hidden 128): second Linear as a plain split GEMM
    size_t oFe1 = (size_t)-1;
    float fe1_inv = 1.f;
    if (oFold16 == (size_t)-1 && proj_umma_supported(C, C, C, false)) {
        oFe1 = h16.size();
        h16.resize(h16.size() + (size_t)2 * C * C);
        fe1_inv = proj_umma_split_weights(hb.data() + oW1, C, C, h16.data() + oFe1);
    }
    // heads: concatenated first layers, zero padded to the accumulator width
    size_t oHeads16 = (size_t)-1;
    int hNpad = 0;
    if (proj_umma_heads_supported(C, h2, nblk, d->num_classes, &hNpad)) {
        std::vector<float> Wp((size_t)hNpad * C, 0.f);
        memcpy(Wp.data(), hW0.data(), sizeof(float) * hW0.size());
        oHeads16 = h16.size();
        h16.resize(h16.size() + (size_t)2 * hNpad * C);
        heads_inv = proj_umma_split_weights(Wp.data(), hNpad, C, h16.data() + oHeads16);
    }

    int prev = 0;
    BGNN_CUDA(cudaGetDevice(&prev));
    BGNN_CUDA(cudaSetDevice(device));
    bgnn_model* m = new bgnn_model();
    m->blob16 = nullptr;
    m->precision = BGNN_PREC_F32; m->blob_bf = nullptr;
    m->opt = default_opts();
    m->d = *d; m->device = device; m->nheads_mlp = nblk;
    cudaError_t e = cudaMalloc(&m->blob, hb.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(m->blob, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !h16.empty()) e = cudaMalloc(&m->blob16, h16.size() * sizeof(uint16_t));
    if (e == cudaSuccess && !h16.empty()) e = cudaMemcpy(m->blob16, h16.data(), h16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    cudaSetDevice(prev);
    if (e != cudaSuccess) { bgnn_set_error("model upload failed: %s", cudaGetErrorString(e)); delete m; return BGNN_E_CUDA; }
    float* B = m->blob;
    m->feW0 = B + oW0; m->feb0 = B + ob0; m->feW1 = B + oW1; m->feb1 = B + ob1;
    m->hc_max = C;
    for (int l = 0; l < d->layers; ++l) {
        LayerDev L;
        layer_shape(d, l, L.in, L.H, L.C, L.HC, L.out);
        L.W = B + lo[l].W; L.att_src = B + lo[l].as; L.att_dst = B + lo[l].ad; L.Medge = B + lo[l].M;
        L.bias = B + lo[l].bias; L.bn_a = B + lo[l].a; L.bn_b = B + lo[l].b;
        L.Wsplit = (o16[l] != (size_t)-1 && m->blob16) ? m->blob16 + o16[l] : nullptr;
        L.w_inv = winv[l];
        L.nblk = 0;
        for (int b = 0; b < 4; ++b) {
            L.Wsplit_blk[b] = (o16blk[l][b] != (size_t)-1 && m->blob16) ? m->blob16 + o16blk[l][b] : nullptr;
            L.w_inv_blk[b] = winvblk[l][b];
            if (L.Wsplit_blk[b]) L.nblk = b + 1;
        }
        L.out_mul = 0.f; L.out_add = 0.f;
        L.Wr = nullptr; L.Wbf = nullptr;
        {   // |bn_a (agg + bias) + bn_b| <= max|bn_a| * max|h| + max(|bn_a||bias| + |bn_b|)
            const float *ba = hb.data() + lo[l].a, *bb = hb.data() + lo[l].b, *bi = hb.data() + lo[l].bias;
            for (int c = 0; c < L.out; ++c) {
                L.out_mul = fmaxf(L.out_mul, fabsf(ba[c]));
                L.out_add = fmaxf(L.out_add, fabsf(ba[c]) * fabsf(bi[c]) + fabsf(bb[c]));
            }
        }
        m->layers.push_back(L);
        if (L.HC > m->hc_max) m->hc_max = L.HC;
    }
    if (nblk * h2 > m->hc_max) m->hc_max = nblk * h2;
    m->hW0 = B + ohW0; m->hb0 = B + ohb0; m->Wcls = B + oWc; m->bcls = B + obc; m->Wconf = B + oWf; m->bconf = B + obf;
    m->Wcorr = rW1 ? B + oWr : nullptr; m->bcorr = rb1 ? B + obr : nullptr;
    m->Wfold0 = (oFold16 != (size_t)-1 && m->blob16) ? m->blob16 + oFold16 : nullptr;
    m->bias_fold0 = B + oFoldBias;
    m->hWsplit = (oHeads16 != (size_t)-1 && m->blob16) ? m->blob16 + oHeads16 : nullptr;
    m->fe1split = (oFe1 != (size_t)-1 && m->blob16) ? m->blob16 + oFe1 : nullptr;
    m->fe1_inv = fe1_inv;
    m->hNpad = hNpad;
    m->fold0_inv = fold0_inv; m->heads_inv = heads_inv;
    m->mlp0_mul = 0.f; m->mlp0_add = 0.f;
    for (int k = 0; k < C; ++k) {       // |relu(x W0^T + b0)| <= max_k sum_f |W0[k,f]| * max|x| + max_k |b0[k]|
        float rs = 0.f;
        for (int f = 0; f < F; ++f) rs += fabsf(hb[oW0 + (size_t)k * F + f]);
        m->mlp0_mul = fmaxf(m->mlp0_mul, rs);
        m->mlp0_add = fmaxf(m->mlp0_add, fabsf(hb[ob0 + k]));
    }
    *out = m;
    return BGNN_OK;
}

extern "C" void bgnn_model_destroy(bgnn_model* m) {
    if (!m) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(m->device);
    cudaFree(m->blob);
    if (m->blob16) cudaFree(m->blob16);
    if (m->blob_bf) cudaFree(m->blob_bf);
    cudaSetDevice(prev);
    delete m;
}

namespace {
__global__ void round_to_bf16(const float* __restrict__ w, float* __restrict__ wr, uint16_t* __restrict__ wbf, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = f32_to_bf16_bits(w[i]);
        wr[i] = bf16_bits_to_f32(b);
        wbf[i] = (uint16_t)b;
    }
}
}  // namespace

extern "C" int bgnn_model_precision(const bgnn_model* m) { return m ? m->precision : -1; }

extern "C" int bgnn_model_set_option(bgnn_model* m, int option, int value) {
    BGNN_CHECK_ARG(m, "null model");
    const bool on = value != 0;
    switch (option) {
        case BGNN_OPT_TENSOR_CORES: m->opt.use_umma = on; break;
        case BGNN_OPT_PRESPLIT: m->opt.presplit = on; break;
        case BGNN_OPT_WINDOW: m->opt.window = on; break;
        case BGNN_OPT_TILE5: m->opt.tile5 = on; break;
        case BGNN_OPT_TILE5_WINDOW: m->opt.tile5_window = on; break;
        case BGNN_OPT_RAGGED_PATCHES: m->opt.ragged_patches = on; break;
        default: BGNN_CHECK_ARG(false, "unknown option %d", option);
    }
    return BGNN_OK;
}

extern "C" int bgnn_model_set_precision(bgnn_model* m, int precision) {
    BGNN_CHECK_ARG(m, "model is null");
    BGNN_CHECK_ARG(precision == BGNN_PREC_F32 || precision == BGNN_PREC_BF16, "unknown precision %d", precision);
    if (precision == BGNN_PREC_BF16 && !m->blob_bf) {     // host-synchronous, once per model
        size_t tot = 0;
        for (const LayerDev& L : m->layers) tot += align_up((size_t)L.HC * L.in, 64);
        int prev = 0;
        BGNN_CUDA(cudaGetDevice(&prev));
        BGNN_CUDA(cudaSetDevice(m->device));
        cudaError_t e = cudaMalloc(&m->blob_bf, tot * (sizeof(float) + sizeof(uint16_t)));
        if (e == cudaSuccess) {
            float* wr = m->blob_bf;
            uint16_t* wb = reinterpret_cast<uint16_t*>(m->blob_bf + tot);
            for (LayerDev& L : m->layers) {
                const int64_t n = (int64_t)L.HC * L.in;
                round_to_bf16<<<(unsigned)ceil_div64(n, 256), 256>>>(L.W, wr, wb, n);
                L.Wr = wr; L.Wbf = wb;
                wr += align_up((size_t)n, 64); wb += align_up((size_t)n, 64);
            }
            e = cudaDeviceSynchronize();
        }
        cudaSetDevice(prev);
        if (e != cudaSuccess) {
            if (m->blob_bf) { cudaFree(m->blob_bf); m->blob_bf = nullptr; }
            for (LayerDev& L : m->layers) { L.Wr = nullptr; L.Wbf = nullptr; }
            bgnn_set_error("bf16 weight preparation failed: %s", cudaGetErrorString(e));
            return BGNN_E_CUDA;
        }
    }
    m->precision = precision;
    return BGNN_OK;
}

static size_t act_bytes(const bgnn_model* m, int64_t n) { return align_up((size_t)n * m->hc_max * sizeof(float), 256); }

extern "C" size_t bgnn_gat_workspace_bytes(const bgnn_model* m, int64_t n_nodes) {
    if (!m || n_nodes < 0) return 0;
    return 3 * act_bytes(m, n_nodes) + align_up((size_t)n_nodes * 2 * m->d.heads * sizeof(float), 256) + 256 /*operand maxima*/ + 256;
}

static bool topo_to_dev(const bgnn_grid_topo* topo, int64_t E, TopoDev& T) {
    if (!(topo && topo->rows > 0 && topo->node_id && topo->n_offsets >= 1 && topo->n_offsets <= 8 && E > 0)) return false;
    T.rows = topo->rows; T.cols = topo->cols; T.node_id = topo->node_id; T.K = topo->n_offsets;
    for (int k = 0; k < 8; ++k) {
        T.dr[k] = k < T.K ? topo->off_dr[k] : 0;
        T.dc[k] = k < T.K ? topo->off_dc[k] : 0;
        if (k < T.K && (abs(T.dr[k]) > 1 || abs(T.dc[k]) > 1)) return false;
    }
    return true;
}

// ragged batch with a band-patch table (bgnn_grid_topo::patches): stencil of the reference's 4 / 8 offsets, dense edge features
static bool ragged_to_dev(const bgnn_model* m, const bgnn_grid_topo* topo, int64_t E, TopoDev& T) {
    if (!(m && m->opt.ragged_patches && topo && topo->rows == 0 && topo->patches && topo->n_patches > 0 && topo->node_id &&
          topo->n_offsets >= 1 && topo->n_offsets <= 8 && E > 0)) return false;
    T.rows = 1; T.cols = 1; T.node_id = topo->node_id; T.K = topo->n_offsets;
    for (int k = 0; k < 8; ++k) {
        T.dr[k] = k < T.K ? topo->off_dr[k] : 0;
        T.dc[k] = k < T.K ? topo->off_dc[k] : 0;
    }
    return true;
}

extern "C" int bgnn_gat_grid_path(const bgnn_model* m, const bgnn_grid_topo* topo, int64_t n_edges) {
    TopoDev T;
    if (m && m->precision == BGNN_PREC_F32 && ragged_to_dev(m, topo, n_edges, T)) {
        for (const LayerDev& L : m->layers)
            if (L.C != 64 || !(L.H == 1 || L.H == 2 || L.H == 4) ||
                !tile_supported(T, L.H, L.C, m->d.edge_dim, topo->grid_eattr, nullptr)) return 0;
        return 1;
    }
    if (!m || m->precision != BGNN_PREC_F32 || !topo_to_dev(topo, n_edges, T)) return 0;
    for (const LayerDev& L : m->layers)
        if (!tile_supported(T, L.H, L.C, m->d.edge_dim, topo->grid_eattr, nullptr)) return 0;
    return 1;
}

namespace {
// n8 groups of 8 consecutive fp32 values -> 8 bf16 (RN)
__global__ void __launch_bounds__(256) f32_to_bf16_rows(const float* __restrict__ in, uint16_t* __restrict__ out, int64_t n8) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
        reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
}
}  // namespace

// bf16 mode (BASELINE configs[4] "bf16 vs fp32 tolerance"; the reference itself has no reduced-precision path).  What is
// bf16 is the STORAGE of the big per-node streams -- the feature-extractor output, the projection weights, h = x W^T and
// every layer output but the last -- i.e. what the aggregation and projection kernels move through HBM; every sum
// (GEMM accumulators, attention logits, softmax, weighted aggregation), BN and the heads stay in fp32.
// oracle/model_oracle.py::forward(bf16_storage=True) restates exactly these rounding points.
static int forward_bf16(const bgnn_model* m, int64_t N, const float* x, const int32_t* rowptr, const int32_t* csr_src,
                        const float* csr_eattr, const Topo5Dev* T5, float* bufA, float* bufB, float* hbuf, float* alpha,
                        float* amax, float* logits,
                        float* probs, int64_t* cls, float* conf, float* corr, int64_t* action, float thr_auto,
                        float thr_review, float* hidden_out, int sms, cudaStream_t s) {
    const bgnn_model_desc& d = m->d;
    const int C = d.hidden, h2 = C / 2;
    int rc;
    { ProfScope ps(10, s);
      if ((rc = launch_gemm(x, d.in_channels, m->feW0, d.in_channels, N, C, bufA, C, m->feb0, 1, C, 0, 0, 0, 0, s))) return rc; }
    { ProfScope ps(11, s);
      if (m->opt.use_umma && m->fe1split && ((int64_t)N * C) % 8 == 0) {
          // second Linear on the tensor cores (fp32-grade split GEMM into hbuf), then one RN pass to the bf16 rows the layer-0
          // projection reads: the same rounding point as the FMA GEMM's bf16 store
          if ((rc = launch_absmax(x, N * d.in_channels, amax, sms, s))) return rc;
          UmmaArgs a = {};
          a.A = bufA; a.M = N; a.K = C; a.Wsplit = m->fe1split; a.w_inv_scale = m->fe1_inv;
          a.amax_src = amax; a.amax_mul = m->mlp0_mul; a.amax_add = m->mlp0_add;
          a.Cout = hbuf; a.bias = m->feb1; a.Htot = 1; a.CH = C; a.head_major = 0;
          if ((rc = launch_umma(a, C, PM_PLAIN, EM_PROJ, sms, s))) return rc;
          const int64_t n8 = (int64_t)N * C / 8;
          int64_t blocks = ceil_div64(n8, 256);
          if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
          f32_to_bf16_rows<<<(unsigned)blocks, 256, 0, s>>>(hbuf, reinterpret_cast<uint16_t*>(bufB), n8);
          BGNN_LAUNCH_CHECK();
      } else if ((rc = launch_gemm(bufA, C, m->feW1, C, N, C, bufB, C, m->feb1, 0, C, 0, 0, 0, 0, s, IO_F32_BF16))) return rc; }
    float* cur = bufB;
    float* nxt = bufA;
    for (int l = 0; l < d.layers; ++l) {
        const LayerDev& L = m->layers[l];
        BGNN_CHECK_ARG(L.Wr, "internal: bf16 weights missing");
        { ProfScope ps(20 + l, s);
          rc = 1;
          if (m->opt.use_umma && proj_bf16_supported(L.HC, L.in, L.C)) {
              rc = launch_proj_bf16(reinterpret_cast<const uint16_t*>(cur), L.Wbf, N, L.in, L.HC, reinterpret_cast<uint16_t*>(hbuf),
                                    L.att_src, L.att_dst, alpha, L.H, L.C, sms, s);
          }
          if (rc == 1)
              rc = launch_gemm(cur, L.in, L.Wr, L.in, N, L.HC, hbuf, L.HC, nullptr, 0, L.C, L.att_src, L.att_dst, alpha, L.H, s, IO_BF16);
          if (rc) return rc; }
        const bool last = l == d.layers - 1;
        { ProfScope ps(40 + l, s);
          const int io = last ? IO_BF16_F32 : IO_BF16;
          rc = 1;
          if (T5) rc = launch_agg_tile5(L, hbuf, alpha, rowptr, csr_eattr, d.edge_dim, !last, d.negative_slope, nxt, N, *T5, io, s);
          if (rc == 1) rc = launch_agg(L, hbuf, alpha, rowptr, csr_src, csr_eattr, d.edge_dim, !last, d.negative_slope, nxt, N, sms, s, io);
          if (rc) return rc; }
        float* t = cur; cur = nxt; nxt = t;
    }
    if (hidden_out) BGNN_CUDA(cudaMemcpyAsync(hidden_out, cur, (size_t)N * C * sizeof(float), cudaMemcpyDeviceToDevice, s));
    const int nblk = m->nheads_mlp;
    if (m->opt.use_umma && m->hWsplit) {         // heads stay fp32: the fused tcgen05 heads kernel of the fp32 path, fed with max|x_L|
        ProfScope ps(60, s);
        BGNN_CUDA(cudaMemsetAsync(amax, 0, 256, s));
        if ((rc = launch_absmax(cur, N * C, amax, sms, s))) return rc;
        UmmaArgs a = {};
        a.A = cur; a.M = N; a.K = C; a.Wsplit = m->hWsplit; a.w_inv_scale = m->heads_inv;
        a.amax_src = amax; a.amax_mul = 1.f; a.amax_add = 0.f;
        a.hb0 = m->hb0; a.h2 = h2; a.nblk = nblk; a.ncls = d.num_classes;
        a.Wcls = m->Wcls; a.bcls = m->bcls; a.Wconf = m->Wconf; a.bconf = m->bconf; a.Wcorr = m->Wcorr; a.bcorr = m->bcorr;
        a.logits = logits; a.probs = probs; a.cls = cls; a.conf = conf; a.corr = corr; a.action = action;
        a.thr_auto = thr_auto; a.thr_review = thr_review;
        return launch_umma(a, m->hNpad, PM_PLAIN, EM_HEADS, sms, s);
    }
    { ProfScope ps(60, s);
      if ((rc = launch_gemm(cur, C, m->hW0, C, N, nblk * h2, hbuf, nblk * h2, m->hb0, 1, C, 0, 0, 0, 0, s))) return rc; }
    int64_t blocks = ceil_div64(N, 8);
    if (blocks > (int64_t)sms * 32) blocks = (int64_t)sms * 32;
    ProfScope ps(61, s);
    heads_finish<<<(unsigned)blocks, 256, 0, s>>>(hbuf, h2, nblk, d.num_classes, m->Wcls, m->bcls, m->Wconf, m->bconf,
                                                 m->Wcorr, m->bcorr, logits, probs, cls, conf, corr, action, thr_auto,
                                                 thr_review, N);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_gat_forward(const bgnn_model* m, int64_t N, int64_t E, const float* x, const int32_t* rowptr,
                                const int32_t* csr_src, const float* csr_eattr, void* workspace, size_t workspace_bytes,
                                float* logits, float* probs, int64_t* cls, float* conf, float* corr, int64_t* action,
                                float thr_auto, float thr_review, float* hidden_out, const bgnn_grid_topo* topo,
                                void* stream_) {
    BGNN_CHECK_ARG(m, "model is null");
    BGNN_CHECK_ARG(N >= 0 && E >= 0, "negative sizes");
    if (N == 0) return BGNN_OK;
    BGNN_CHECK_ARG(x, "x is null");
    const bool grid_only = bgnn_gat_grid_path(m, topo, E) == 1;       // every layer tiled: the CSR payload is not read
    BGNN_CHECK_ARG(grid_only || rowptr, "rowptr is null");
    BGNN_CHECK_ARG(grid_only || E == 0 || csr_src, "csr_src is null");
    BGNN_CHECK_ARG(grid_only || m->d.edge_dim == 0 || E == 0 || csr_eattr, "model has edge_dim>0 but csr_eattr is null");
    BGNN_CHECK_ARG(N < (int64_t)2000000000 && E < (int64_t)2000000000, "graph too large for int32 indices");
    if (workspace_bytes < bgnn_gat_workspace_bytes(m, N) || !workspace) {
        bgnn_set_error("workspace too small: %zu < %zu", workspace_bytes, bgnn_gat_workspace_bytes(m, N));
        return BGNN_E_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream_;
    int dev = 0, sms = 148;
    BGNN_CUDA(cudaGetDevice(&dev));
    BGNN_CHECK_ARG(dev == m->device, "model lives on device %d but current device is %d", m->device, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

    // grid path: radius-1 stencil (4-/8-connected) on a single tile
    TopoDev T;
    const bool tiled = topo_to_dev(topo, E, T);
    TopoDev TR;
    const bool ragged = !tiled && ragged_to_dev(m, topo, E, TR);
    Topo5Dev T5;
    const bool tiled5 = m->opt.tile5 && topo5_to_dev(topo, E, T5);
    T5.window = m->opt.tile5_window;

    char* w = (char*)workspace;
    w = (char*)align_up((size_t)w, 256);
    float* bufA = (float*)w; w += act_bytes(m, N);
    float* bufB = (float*)w; w += act_bytes(m, N);
    float* hbuf = (float*)w; w += act_bytes(m, N);
    float* alpha = (float*)w; w += align_up((size_t)N * 2 * m->d.heads * sizeof(float), 256);
    // operand maxima for the fp16-split tensor-core GEMMs: [0] = max|x| (or max|MLP output|), [1+l] = max|h_l|
    float* amax = (float*)w; w += 256;
    const bgnn_model_desc& d = m->d;
    BGNN_CHECK_ARG(d.layers + 1 <= 64, "too many layers");
    if (m->opt.use_umma) BGNN_CUDA(cudaMemsetAsync(amax, 0, 256, s));
    const int C = d.hidden, h2 = C / 2;
    int rc;
    if (m->precision == BGNN_PREC_BF16)
        return forward_bf16(m, N, x, rowptr, csr_src, csr_eattr, tiled5 ? &T5 : nullptr, bufA, bufB, hbuf, alpha, amax, logits, probs, cls, conf, corr, action,
                            thr_auto, thr_review, hidden_out, sms, s);
    // LocalFeatureExtractor: Linear(F,C) + ReLU, Linear(C,C).  On the tensor-core path both are fused into the
    // layer-0 projection (first Linear computed by its producers, second folded into its weights).
    const bool fused0 = m->opt.use_umma && m->Wfold0;
    if (!fused0) {
        { ProfScope ps(10, s);
          if ((rc = launch_gemm(x, d.in_channels, m->feW0, d.in_channels, N, C, bufA, C, m->feb0, 1, C, 0, 0, 0, 0, s))) return rc; }
        { ProfScope ps(11, s);
          if (m->opt.use_umma && m->fe1split) {      // |relu(x W0^T + b0)| <= mlp0_mul * max|x| + mlp0_add bounds the operand
              if ((rc = launch_absmax(x, N * d.in_channels, amax, sms, s))) return rc;
              UmmaArgs a = {};
              a.A = bufA; a.M = N; a.K = C; a.Wsplit = m->fe1split; a.w_inv_scale = m->fe1_inv;
              a.amax_src = amax; a.amax_mul = m->mlp0_mul; a.amax_add = m->mlp0_add;
              a.Cout = bufB; a.bias = m->feb1; a.Htot = 1; a.CH = C; a.head_major = 0;
              if ((rc = launch_umma(a, C, PM_PLAIN, EM_PROJ, sms, s))) return rc;
              BGNN_CUDA(cudaMemsetAsync(amax, 0, sizeof(float), s));     // slot 0 is re-measured on the MLP output below
          } else if ((rc = launch_gemm(bufA, C, m->feW1, C, N, C, bufB, C, m->feb1, 0, C, 0, 0, 0, 0, s))) return rc; }
    }
    float* cur = bufB;
    float* nxt = bufA;
    const float* hsrc = hbuf;                  // where h_l and its alpha live
    const float* al = alpha;
    bool cur_split = false;                    // `cur` holds fp16 hi / lo planes [N, in] (written by the tiled aggregation)
    bool is4_ = false;
    const bool tile_ok0 = tiled && m->opt.use_umma && m->opt.presplit;
    // no single grid (ragged batch): a topology with rows == 0 only carries the stencil size = an in-degree bound, which
    // lets the CSR aggregation run as the windowed shared-memory kernel
    const bool windowed = !tiled && m->opt.window && topo && topo->rows == 0 && topo->n_offsets >= 1 && topo->n_offsets <= 8 && E > 0;
    for (int l = 0; l < d.layers; ++l) {
        const LayerDev& L = m->layers[l];
        { ProfScope ps(20 + l, s);
          if (l == 0 && fused0) {
              if ((rc = launch_absmax(x, N * d.in_channels, amax, sms, s))) return rc;
              UmmaArgs a = {};
              a.M = N; a.K = C; a.Wsplit = m->Wfold0; a.w_inv_scale = m->fold0_inv;
              a.amax_src = amax; a.amax_mul = m->mlp0_mul; a.amax_add = m->mlp0_add; a.cmax_out = amax + 1;
              a.x = x; a.F = d.in_channels; a.W0 = m->feW0; a.b0 = m->feb0;
              a.Cout = hbuf; a.bias = m->bias_fold0; a.att_src = L.att_src; a.att_dst = L.att_dst; a.alpha = alpha;
              a.Htot = L.H; a.CH = L.C; a.head_major = 1;
              rc = launch_umma(a, L.HC, PM_MLP0, EM_PROJ, sms, s);
          } else if (m->opt.use_umma && L.Wsplit) {
              UmmaArgs a = {};
              a.A = cur; a.M = N; a.K = L.in; a.Wsplit = L.Wsplit; a.w_inv_scale = L.w_inv;
              if (cur_split) { a.A_hi = reinterpret_cast<const uint16_t*>(cur); a.A_lo = a.A_hi + L.in; a.a_ld = 2 * L.in; }
              if (l == 0) {               // input = feature-extractor output (unfused): measure it
                  if ((rc = launch_absmax(cur, N * L.in, amax, sms, s))) return rc;
                  a.amax_src = amax; a.amax_mul = 1.f; a.amax_add = 0.f;
              } else {
                  a.amax_src = amax + l; a.amax_mul = m->layers[l - 1].out_mul; a.amax_add = m->layers[l - 1].out_add;
              }
              a.cmax_out = amax + 1 + l;
              a.Cout = hbuf; a.att_src = L.att_src; a.att_dst = L.att_dst; a.alpha = alpha; a.Htot = L.H; a.CH = L.C;
              a.head_major = 1;
              rc = launch_umma(a, L.HC, cur_split ? PM_SPLIT : PM_PLAIN, EM_PROJ, sms, s);
          } else if (m->opt.use_umma && L.nblk == 4) {      // H*C = 512: four column blocks of 128 (whole heads), one launch each
              if (l == 0 && (rc = launch_absmax(cur, N * L.in, amax, sms, s))) return rc;
              const int hpb = 128 / L.C;               // heads per block
              for (int b = 0; b < L.nblk; ++b) {
                  UmmaArgs a = {};
                  a.A = cur; a.M = N; a.K = L.in; a.Wsplit = L.Wsplit_blk[b]; a.w_inv_scale = L.w_inv_blk[b];
                  if (l == 0) { a.amax_src = amax; a.amax_mul = 1.f; a.amax_add = 0.f; }
                  else { a.amax_src = amax + l; a.amax_mul = m->layers[l - 1].out_mul; a.amax_add = m->layers[l - 1].out_add; }
                  a.cmax_out = amax + 1 + l;
                  a.Cout = hbuf + (size_t)b * N * 128;                     // head-major planes of this block's heads
                  a.att_src = L.att_src + b * 128; a.att_dst = L.att_dst + b * 128;
                  a.alpha = alpha + b * hpb; a.Htot = L.H; a.CH = L.C; a.head_major = 1;
                  if ((rc = launch_umma(a, 128, PM_PLAIN, EM_PROJ, sms, s))) break;
              }
          } else {
              rc = launch_gemm(cur, L.in, L.W, L.in, N, L.HC, hbuf, L.HC, nullptr, 0, L.C, L.att_src, L.att_dst, alpha, L.H, s);
              if (!rc && m->opt.use_umma) rc = launch_absmax(hbuf, N * L.HC, amax + 1 + l, sms, s);   // a later tensor-core GEMM needs max|h_l|
          }
          if (rc) return rc; }
        const int relu = l < d.layers - 1;
        { ProfScope ps(40 + l, s);
          rc = 1;
          // when the next layer's projection runs on the tensor cores, hand it the output as fp16 hi / lo planes
          SplitOut SO = {};
          bool want_split = false;
          const bool win_l = windowed && L.C == 64 && (L.H == 1 || L.H == 2 || L.H == 4);
          const bool rag_l = ragged && L.C == 64 && (L.H == 1 || L.H == 2 || L.H == 4) &&
                             tile_supported(TR, L.H, L.C, d.edge_dim, topo->grid_eattr, nullptr);
          if ((tile_ok0 || ((win_l || rag_l) && m->opt.use_umma && m->opt.presplit)) && l + 1 < d.layers && L.C % 64 == 0) {
              const LayerDev& Ln = m->layers[l + 1];
              want_split = Ln.Wsplit && Ln.in == L.HC && proj_umma_presplit_supported(Ln.HC, Ln.in, Ln.C, true) &&
                           (win_l || rag_l || tile_supported(T, L.H, L.C, d.edge_dim, topo->grid_eattr, &is4_));
          }
          // last layer: the heads' tensor-core kernel takes the same hand-over (no converter stage; bit-identical products) unless
          // the caller wants the fp32 hidden activations back
          if ((tile_ok0 || ((win_l || rag_l) && m->opt.use_umma && m->opt.presplit)) && l + 1 == d.layers && L.HC == 64 && !hidden_out &&
              m->hWsplit && proj_umma_heads_presplit_supported(C, m->hNpad))
              want_split = win_l || rag_l || tile_supported(T, L.H, L.C, d.edge_dim, topo->grid_eattr, &is4_);
          if (want_split) {
              SO.hi = reinterpret_cast<uint16_t*>(nxt);
              SO.amax_src = amax + 1 + l; SO.mul = L.out_mul; SO.add = L.out_add;
          }
          cur_split = false;
          if (rc == 1 && tiled) {
              rc = launch_agg_tile(L, hsrc, al, T, topo->grid_eattr, d.edge_dim, relu, d.negative_slope, nxt, N, SO, s);
              if (rc == 0 && want_split) cur_split = true;
          }
          if (rc == 1 && rag_l) {
              rc = launch_agg_ragged(L, hsrc, al, topo->node_id, topo->patches, topo->n_patches, TR, topo->grid_eattr, d.edge_dim, relu,
                                     d.negative_slope, nxt, N, SO, s);
              if (rc == 0 && want_split) cur_split = true;
          }
          if (rc == 1 && win_l) {
              rc = launch_agg_window(L, hsrc, al, rowptr, csr_src, csr_eattr, d.edge_dim, relu, d.negative_slope, nxt, N, SO, s);
              if (rc == 0 && want_split) cur_split = true;
          }
          if (rc == 1 && tiled5 && !tiled)        // 5 x 5 stencil on a single grid: radius-2 tiled kernel
              rc = launch_agg_tile5(L, hsrc, al, rowptr, csr_eattr, d.edge_dim, relu, d.negative_slope, nxt, N, T5, IO_F32, s);
          if (rc == 1) rc = launch_agg(L, hsrc, al, rowptr, csr_src, csr_eattr, d.edge_dim, relu, d.negative_slope, nxt, N, sms, s);
          if (rc) return rc; }
        float* t = cur; cur = nxt; nxt = t;
    }
    if (hidden_out) BGNN_CUDA(cudaMemcpyAsync(hidden_out, cur, (size_t)N * C * sizeof(float), cudaMemcpyDeviceToDevice, s));
    // heads
    const int nblk = m->nheads_mlp;
    if (m->opt.use_umma && m->hWsplit) {         // one kernel: first layers on the tensor cores, the rest in its epilogue
        ProfScope ps(60, s);
        UmmaArgs a = {};
        a.A = cur; a.M = N; a.K = C; a.Wsplit = m->hWsplit; a.w_inv_scale = m->heads_inv;
        if (cur_split) { a.A_hi = reinterpret_cast<const uint16_t*>(cur); a.A_lo = a.A_hi + C; a.a_ld = 2 * C; }
        const LayerDev& LL = m->layers[d.layers - 1];
        a.amax_src = amax + d.layers; a.amax_mul = LL.out_mul; a.amax_add = LL.out_add;
        a.hb0 = m->hb0; a.h2 = h2; a.nblk = nblk; a.ncls = d.num_classes;
        a.Wcls = m->Wcls; a.bcls = m->bcls; a.Wconf = m->Wconf; a.bconf = m->bconf; a.Wcorr = m->Wcorr; a.bcorr = m->bcorr;
        a.logits = logits; a.probs = probs; a.cls = cls; a.conf = conf; a.corr = corr; a.action = action;
        a.thr_auto = thr_auto; a.thr_review = thr_review;
        return launch_umma(a, m->hNpad, cur_split ? PM_SPLIT : PM_PLAIN, EM_HEADS, sms, s);
    }
    { ProfScope ps(60, s);
      if ((rc = launch_gemm(cur, C, m->hW0, C, N, nblk * h2, hbuf, nblk * h2, m->hb0, 1, C, 0, 0, 0, 0, s))) return rc; }
    int64_t blocks = ceil_div64(N, 8);
    if (blocks > (int64_t)sms * 32) blocks = (int64_t)sms * 32;
    ProfScope ps(61, s);
    heads_finish<<<(unsigned)blocks, 256, 0, s>>>(hbuf, h2, nblk, d.num_classes, m->Wcls, m->bcls, m->Wconf, m->bconf,
                                                 m->Wcorr, m->bcorr, logits, probs, cls, conf, corr, action, thr_auto,
                                                 thr_review, N);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

// ------------------------------------------------------------------------------------------ GCN / GraphSAGE / GIN backbones
// models/gnn.py:120-143 (`gnn_type != "GAT"`): layers without attention on the same in-edge CSR.  The host layer
// (bathymetric-gnn_b200/gnn.py::_run_plain) composes a layer from these entry points; PyG semantics as restated in
// oracle/pyg_restated.py (GCNConv / SAGEConv / GINConv).  Hidden width only (64 by default): 256 B per node and layer, a small
// fraction of the GAT path's traffic, so one warp-per-destination gather kernel + the fp32 FMA GEMM serve all three.
namespace {
// one warp per destination; lane <-> channels (strided); sources in CSR order = the reference's edge order, self term last
template <int MODE>
__global__ void __launch_bounds__(256)
csr_reduce(int64_t N, int C, const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
           const int32_t* __restrict__ src, const float* __restrict__ dis_tab, int dis_n, float eps,
           const float* __restrict__ pre_add, const float* __restrict__ mul, const float* __restrict__ add, int relu,
           float* __restrict__ out, int64_t ldo, int copy_self) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = wid; i < N; i += nw) {
        const int e0 = __ldg(rowptr + i), e1 = __ldg(rowptr + i + 1);
        const int deg = e1 - e0;
        float di = 0.f;
        if (MODE == BGNN_AGG_GCN) di = __ldg(dis_tab + min(deg + 1, dis_n - 1));          // deg^-1/2 of (in-degree + self loop)
        for (int c = lane; c < C; c += 32) {
            float acc = 0.f;
            for (int e = e0; e < e1; ++e) {
                const int j = __ldg(src + e);
                const float v = __ldg(x + (int64_t)j * ldx + c);
                if (MODE == BGNN_AGG_GCN) {
                    const int dj = __ldg(rowptr + j + 1) - __ldg(rowptr + j);
                    const float nrm = __ldg(dis_tab + min(dj + 1, dis_n - 1)) * di;       // (dis[src] * 1) * dis[dst]   (gcn_norm)
                    acc += nrm * v;
                } else {
                    acc += v;
                }
            }
            const float xi = __ldg(x + i * ldx + c);
            float y;
            if (MODE == BGNN_AGG_GCN) y = acc + (di * di) * xi;                           // self loop appended behind the edges
            else if (MODE == BGNN_AGG_MEAN) y = acc / (float)(deg > 0 ? deg : 1);         // scatter-mean: sum / clamp(count, 1)
            else y = acc + (1.0f + eps) * xi;                                             // GIN: out + (1 + eps) * x
            if (pre_add) y += __ldg(pre_add + c);
            if (mul) y = y * __ldg(mul + c) + (add ? __ldg(add + c) : 0.f);
            if (relu) y = fmaxf(y, 0.f);
            out[i * ldo + c] = y;
            if (copy_self) out[i * ldo + C + c] = xi;                                     // SAGE: [mean | x] rows for ONE GEMM
        }
    }
}
}  // namespace

extern "C" int bgnn_csr_aggregate(int mode, int64_t n_nodes, int channels, const float* x, int64_t ldx, const int32_t* rowptr,
                                  const int32_t* csr_src, const float* deg_inv_sqrt, int deg_table_len, float eps,
                                  const float* pre_add, const float* mul, const float* add, int relu, float* out, int64_t ldo,
                                  int copy_self, void* stream) {
    BGNN_CHECK_ARG(n_nodes >= 0 && channels > 0, "bad shape");
    if (n_nodes == 0) return BGNN_OK;
    BGNN_CHECK_ARG(x && rowptr && out && ldx >= channels && ldo >= channels * (copy_self ? 2 : 1), "null / short argument");
    BGNN_CHECK_ARG(mode != BGNN_AGG_GCN || (deg_inv_sqrt && deg_table_len >= 2), "GCN needs the deg^-1/2 table");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(80, s);
    const unsigned blocks = (unsigned)(ceil_div64(n_nodes, 8) > 148 * 32 ? 148 * 32 : ceil_div64(n_nodes, 8));
#define RED(M_) csr_reduce<M_><<<blocks, 256, 0, s>>>(n_nodes, channels, x, ldx, rowptr, csr_src, deg_inv_sqrt, deg_table_len, eps, pre_add, mul, add, relu, out, ldo, copy_self)
    if (mode == BGNN_AGG_GCN) RED(BGNN_AGG_GCN);
    else if (mode == BGNN_AGG_MEAN) RED(BGNN_AGG_MEAN);
    else if (mode == BGNN_AGG_GIN) RED(BGNN_AGG_GIN);
    else BGNN_CHECK_ARG(false, "unknown aggregation mode %d", mode);
#undef RED
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

extern "C" int bgnn_linear(const float* x, int64_t ldx, int64_t n_rows, int in_features, const float* weight, const float* bias,
                           int relu, int out_features, float* out, int64_t ldo, void* stream) {
    BGNN_CHECK_ARG(n_rows >= 0 && in_features > 0 && out_features > 0, "bad shape");
    if (n_rows == 0) return BGNN_OK;
    BGNN_CHECK_ARG(x && weight && out && ldx >= in_features && ldo >= out_features && ldx < (1ll << 31) && ldo < (1ll << 31), "null / short argument");
    ProfScope ps(81, (cudaStream_t)stream);
    return launch_gemm(x, (int)ldx, weight, in_features, n_rows, out_features, out, (int)ldo, bias, relu, 64, nullptr, nullptr,
                       nullptr, 1, (cudaStream_t)stream);
}

extern "C" int bgnn_heads_finish(const float* hidden, int64_t n_nodes, int h2, int nblk, int num_classes, const float* Wcls,
                                 const float* bcls, const float* Wconf, const float* bconf, const float* Wcorr, const float* bcorr,
                                 float* logits, float* probs, int64_t* cls, float* conf, float* corr, int64_t* action,
                                 float thr_auto, float thr_review, void* stream) {
    BGNN_CHECK_ARG(n_nodes >= 0 && h2 > 0 && (nblk == 2 || nblk == 3) && num_classes >= 1 && num_classes <= 8, "bad shape");
    if (n_nodes == 0) return BGNN_OK;
    BGNN_CHECK_ARG(hidden && Wcls && bcls && Wconf && bconf && (nblk == 2 || (Wcorr && bcorr)), "null argument");
    ProfScope ps(61, (cudaStream_t)stream);
    const int64_t blocks = ceil_div64(n_nodes, 8) > 148 * 16 ? 148 * 16 : ceil_div64(n_nodes, 8);
    heads_finish<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(hidden, h2, nblk, num_classes, Wcls, bcls, Wconf, bconf, Wcorr, bcorr,
                                                                     logits, probs, cls, conf, corr, action, thr_auto, thr_review, n_nodes);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

#ifdef BGNN_AGG_TIMING
extern "C" int bgnn_debug_agg_timing(unsigned long long* out8, int reset) {
    if (out8) cudaMemcpyFromSymbol(out8, g_agg_t, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(g_agg_t, z, sizeof(z)); }
    return 0;
}
#endif
