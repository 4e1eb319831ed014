// tcgen05 projection GEMM (see proj_umma.cu).  Internal to libbgnn.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

enum { PM_PLAIN = 0, PM_MLP0 = 1, PM_SPLIT = 2 };   // producer: fp32 A from global memory | A = relu(x W0^T + b0) computed on
                                                    // the fly | A already split into fp16 hi / lo planes by the previous kernel
enum { EM_PROJ = 0, EM_HEADS = 1 };      // epilogue: C (+bias) and alpha_src/alpha_dst | finish the three output heads

struct UmmaArgs {
    // operands: C[M,N] = A[M,K] . W[N,K]^T, W pre-split (2 x fp16, scaled by 1/w_inv_scale) and pre-swizzled on the host
    const float* A; int64_t M; int K; const uint16_t* Wsplit; float w_inv_scale;
    // PM_SPLIT: A_hi / A_lo [M,K] fp16 bit patterns of a*sA (sA from the same amax bound below), written by the aggregation
    const uint16_t* A_hi; const uint16_t* A_lo; int64_t a_ld;     // a_ld: row stride in elements (0 = K)
    // upper bound of max|A| = amax_mul * (*amax_src) + amax_add (amax_src may be NULL) -> power-of-two operand scale
    const float* amax_src; float amax_mul, amax_add;
    float* cmax_out;          // EM_PROJ, optional: atomicMax of |C| over everything this launch writes
    // PM_MLP0: A[m,k] = relu(b0[k] + sum_f x[m,f] W0[k,f]), x [M,F] row-major, F <= 8, K <= 128
    const float* x; int F; const float* W0; const float* b0;
    // EM_PROJ: Cout [M,N] (+ bias [N], optional); alpha (optional) [M, 2*Htot] from att_src / att_dst [N]
    float* Cout; const float* bias; const float* att_src; const float* att_dst; float* alpha; int Htot; int CH;
    int head_major;           // EM_PROJ: store C as [Htot][M][CH] (one contiguous plane per head) instead of [M][N]
    int mc;                   // set by launch_umma: the CTA pair of an N = 256 PM_SPLIT launch runs as a 2-CTA cluster and each CTA
                              // multicasts one fp16 plane of every A chunk to both (A crosses HBM / L2 once per pair)
    // EM_HEADS: W = concatenated first layers [nblk*h2 (zero padded to N), K]; hb0 their biases; second layers below
    const float* hb0; int h2, nblk, ncls;
    const float *Wcls, *bcls, *Wconf, *bconf, *Wcorr, *bcorr;
    float* logits; float* probs; int64_t* cls; float* conf; float* corr; int64_t* action;
    float thr_auto, thr_review;
};

// shapes the tensor-core path is compiled for (N in {32,64,128,256}, K multiple of 32)
bool proj_umma_supported(int N, int K, int CH, bool want_alpha);
bool proj_umma_presplit_supported(int N, int K, int CH, bool want_alpha);   // PM_SPLIT + EM_PROJ
bool proj_umma_mlp0_supported(int N, int K, int F);               // PM_MLP0 + EM_PROJ
bool proj_umma_heads_supported(int K, int h2, int nblk, int ncls, int* Npad);   // PM_PLAIN + EM_HEADS
bool proj_umma_heads_presplit_supported(int K, int Npad);                       // PM_SPLIT + EM_HEADS (A handed over as fp16 hi / lo planes)
// host: W [N,K] fp32 -> 2 x fp16 split of W*s, [K/32][2][N x 32 swizzled tile] (2*N*K uint16); returns 1/s
float proj_umma_split_weights(const float* W, int N, int K, uint16_t* out);
// max |v| over n floats, atomicMax into *out (float bits; zero it first)
int launch_absmax(const float* v, int64_t n, float* out, int sms, cudaStream_t s);
// general entry: N = accumulator width (for EM_HEADS the padded width)
int launch_umma(const UmmaArgs& a, int N, int pmode, int emode, int sms, cudaStream_t s);

// ---- bf16 mode (proj_bf16.cu): C = A . W^T with bf16 operands straight from HBM (one tcgen05 product, fp32 accumulation in
// TMEM), C stored as bf16 head-major [Htot][M][CH], alpha_src / alpha_dst from the fp32 accumulators.
bool proj_bf16_supported(int N, int K, int CH);
int launch_proj_bf16(const uint16_t* A, const uint16_t* W, int64_t M, int K, int N, uint16_t* Cout, const float* att_src,
                     const float* att_dst, float* alpha, int Htot, int CH, int sms, cudaStream_t s);
