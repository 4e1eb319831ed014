// tcgen05 projection GEMM (see proj_umma.cu).  Internal to libbgnn.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

// shapes the tensor-core path is compiled for (N in {32,64,128,256}, K multiple of 32)
bool proj_umma_supported(int N, int K, int CH, bool want_alpha);
// host: W [N,K] fp32 -> three bf16 matrices [3][N][K] with W = w1 + w2 + w3 (round-to-nearest splits)
void proj_umma_split_weights(const float* W, int N, int K, uint16_t* out);
// C[M,N] = A[M,K] . W^T; alpha (optional) [M, 2*Htot]: per-head dot products with att_src / att_dst
int launch_proj_umma(const float* A, int64_t M, int K, const uint16_t* Wsplit, int N, float* C, const float* att_src,
                     const float* att_dst, float* alpha, int Htot, int CH, int sms, cudaStream_t s);
