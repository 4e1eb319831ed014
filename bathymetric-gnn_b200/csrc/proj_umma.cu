// Projection GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//     C[M, N] = A[M, K] . W[N, K]^T      (fp32 in, fp32 out)     + optional alpha epilogue
//
// This is the `h = x W^T` of every GATConv layer (PyG GATConv.lin, called from models/gnn.py:175-176),
// a skinny dense contraction: M = nodes (millions), N, K <= 256.  fp32 parity (atol 2e-5 on logits)
// rules out plain bf16 / tf32 operands, so both operands are split into three bf16 terms
// (a = a1 + a2 + a3, exact to 24 bits) and the six products with i + j <= 4 are accumulated in
// fp32 in TMEM: error ~2^-23 per product, i.e. fp32-grade, at 1/6 of the bf16 tensor rate --
// still ~8x the fp32 FMA pipe.
//
// The tensor core accumulates in fp32 with truncation (RZ), so a long accumulation chain biases the
// result.  Following Ootomo & Yokota (2022) the leading product a1*b1 gets its OWN accumulator (K/16
// chained MMAs only) and the five correction products go to a second accumulator that is 2^-8 smaller in
// magnitude; the two are added in fp32 (RN) in the epilogue.
//
// One persistent CTA per SM, 256 threads.  A "super tile" is MT x 128 rows (MT = 1 for N = 256, else 2);
// TMEM holds {main, corr} x MT x N <= 512 columns.
//   producers (all 8 warps): A chunk [MT*128 x 32] fp32 -> LDG.128 -> 3-way bf16 split -> STS into
//                            K-major SWIZZLE_64B tiles; W splits (pre-split on the host) via cp.async
//   MMA issue (one thread):  12 x tcgen05.mma.kind::f16 (M=128, N, K=16) per M tile and chunk,
//                            tcgen05.commit -> mbarrier frees the stage (3-stage ring)
//   epilogue (warp w -> TMEM lanes 32*(w%4)..): tcgen05.ld 32x32b.x32 of main and corr, fp32 add,
//                            alpha_src / alpha_dst dot products per head, fp32 rows to global
#include <cuda_bf16.h>

#include "common.cuh"
#include "proj_umma.cuh"

namespace {

constexpr int UM_THREADS = 416;           // 4 epilogue warps + 1 MMA warp + 8 producer warps
constexpr int KC = 32;                 // K elements per stage (64 B of bf16 = one SWIZZLE_64B row)
constexpr int STAGES = 3;
constexpr int TILE_M = 128;
constexpr int A_TILE_BYTES = TILE_M * KC * 2;     // 8 KB per split
constexpr uint32_t SBO_BYTES = 8 * KC * 2;        // 512 B between 8-row groups

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cp_async16(uint32_t sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdst), "l"(gsrc) : "memory");
}

// K-major SWIZZLE_64B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) | version=1 [46,48) | layout=4 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((SBO_BYTES >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of the 16 B chunk holding elements [8c, 8c+8) of row r inside a K-major SWIZZLE_64B tile
__device__ __forceinline__ uint32_t sw64_off(int r, int c) {
    return (uint32_t)((r >> 3) * SBO_BYTES + (r & 7) * 64 + ((c ^ ((r & 7) >> 1)) << 4));
}

struct Split3 { __nv_bfloat162 p[3][2]; };     // [split][pair] for 4 consecutive fp32
__device__ __forceinline__ void split4(const float4 v, uint2& s1, uint2& s2, uint2& s3) {
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 b1[4], b2[4], b3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        b1[i] = __float2bfloat16_rn(x[i]);
        const float r1 = x[i] - __bfloat162float(b1[i]);
        b2[i] = __float2bfloat16_rn(r1);
        const float r2 = r1 - __bfloat162float(b2[i]);
        b3[i] = __float2bfloat16_rn(r2);
    }
    auto pack = [](const __nv_bfloat16* b) {
        uint2 o;
        o.x = (uint32_t)__bfloat16_as_ushort(b[0]) | ((uint32_t)__bfloat16_as_ushort(b[1]) << 16);
        o.y = (uint32_t)__bfloat16_as_ushort(b[2]) | ((uint32_t)__bfloat16_as_ushort(b[3]) << 16);
        return o;
    };
    s1 = pack(b1); s2 = pack(b2); s3 = pack(b3);
}

template <int N>
struct UmmaSmem {
    static constexpr int MT = N >= 256 ? 1 : 2;                  // M tiles per super tile
    static constexpr int SUPER_M = MT * TILE_M;
    static constexpr int B_TILE_BYTES = N * KC * 2;
    static constexpr int STAGE_BYTES = MT * 3 * A_TILE_BYTES + 3 * B_TILE_BYTES;
    static constexpr int ACC_COLS = 2 * MT * N;                  // {main, corr} per M tile
    static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
    static constexpr size_t BYTES = (size_t)STAGES * STAGE_BYTES + 2 * N * sizeof(float) + 128;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_global_v8(float* p, const float* x) {          // 256-bit store (PTX ISA 8.8, sm_100)
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]),
                 "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]) : "memory");
}

// warp roles
constexpr int EPI_WARPS = 4, PROD_WARPS = 8, PROD_GROUPS = 2;
constexpr int MMA_WARP = EPI_WARPS, PROD_WARP0 = EPI_WARPS + 1;
constexpr int PROD_GROUP_THREADS = PROD_WARPS / PROD_GROUPS * 32;                  // 128
static_assert(UM_THREADS == (EPI_WARPS + 1 + PROD_WARPS) * 32, "thread count / role mismatch");
// barrier slots
constexpr int BAR_FULL = 0, BAR_EMPTY = STAGES, BAR_TFULL = 2 * STAGES, BAR_TEMPTY = 2 * STAGES + 1, NUM_BARS = 2 * STAGES + 2;

template <int N>
__global__ void __launch_bounds__(UM_THREADS, 1)
proj_umma_kernel(const float* __restrict__ A, int64_t M, int K, const uint16_t* __restrict__ Wsplit /*[3][N][K] bf16*/,
                 float* __restrict__ Cout, const float* __restrict__ att_src, const float* __restrict__ att_dst,
                 float* __restrict__ alpha, int Htot, int CH) {
    using S = UmmaSmem<N>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw;                                                      // stage ring (1024 B aligned)
    float* s_att = reinterpret_cast<float*>(base + (size_t)STAGES * S::STAGE_BYTES);      // [2][N]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_att + 2 * N);                          // NUM_BARS
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

    constexpr int MT = S::MT, SUPER_M = S::SUPER_M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkc = K / KC;
    const int64_t n_super = (M + SUPER_M - 1) / SUPER_M;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bars[BAR_FULL + s], PROD_GROUP_THREADS); mbar_init(&bars[BAR_EMPTY + s], 1); }
        mbar_init(&bars[BAR_TFULL], 1);
        mbar_init(&bars[BAR_TEMPTY], EPI_WARPS * 32);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (alpha) {
        for (int i = tid; i < N; i += UM_THREADS) { s_att[i] = __ldg(att_src + i); s_att[N + i] = __ldg(att_dst + i); }
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "n"(S::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    if (warp >= PROD_WARP0) {
        // ===================== PRODUCERS: two groups of 64 threads own alternate K chunks =====================
        const int grp = (warp - PROD_WARP0) / (PROD_WARPS / PROD_GROUPS);
        const int ptid = tid - (PROD_WARP0 + grp * (PROD_WARPS / PROD_GROUPS)) * 32;      // 0..63
        constexpr int A_F4 = MT * TILE_M * (KC / 4) / PROD_GROUP_THREADS;                  // 8 or 16 float4 per thread and chunk
        uint32_t it = 0;
        for (int64_t st = blockIdx.x; st < n_super; st += gridDim.x) {
            const int64_t row0 = st * SUPER_M;
            for (int kc = 0; kc < nkc; ++kc, ++it) {
                if ((int)(it % PROD_GROUPS) != grp) continue;
                const int s = it % STAGES;
                unsigned char* stage = base + (size_t)s * S::STAGE_BYTES;
                // A loads first (long latency, registers only), then wait for the stage to be free
                float4 av[A_F4];
#pragma unroll
                for (int i = 0; i < A_F4; ++i) {
                    const int f = ptid + i * PROD_GROUP_THREADS;
                    const int mt = f >> 10, rem = f & 1023, r = rem >> 3, c4 = rem & 7;
                    const int64_t gr = row0 + mt * TILE_M + r;
                    av[i] = gr < M ? __ldg(reinterpret_cast<const float4*>(A + gr * K + kc * KC + c4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (it >= STAGES) mbar_wait(&bars[BAR_EMPTY + s], ((it / STAGES) - 1) & 1);
                {   // W split chunks: 3 x N rows x 64 B, straight to shared memory
                    const uint32_t sB = smem_u32(stage + MT * 3 * A_TILE_BYTES);
                    for (int q = ptid; q < 3 * N * 4; q += PROD_GROUP_THREADS) {
                        const int sp = q / (N * 4), rem = q - sp * (N * 4), n = rem >> 2, c = rem & 3;
                        cp_async16(sB + sp * S::B_TILE_BYTES + sw64_off(n, c), Wsplit + ((size_t)sp * N + n) * K + kc * KC + c * 8);
                    }
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
                }
#pragma unroll
                for (int i = 0; i < A_F4; ++i) {                   // 3-way bf16 split -> swizzled STS
                    const int f = ptid + i * PROD_GROUP_THREADS;
                    const int mt = f >> 10, rem = f & 1023, r = rem >> 3, c4 = rem & 7;
                    uint2 s1, s2, s3;
                    split4(av[i], s1, s2, s3);
                    unsigned char* t0 = stage + (size_t)(mt * 3) * A_TILE_BYTES + sw64_off(r, c4 >> 1) + (c4 & 1) * 8;
                    *reinterpret_cast<uint2*>(t0) = s1;
                    *reinterpret_cast<uint2*>(t0 + A_TILE_BYTES) = s2;
                    *reinterpret_cast<uint2*>(t0 + 2 * A_TILE_BYTES) = s3;
                }
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy writes -> async proxy (tensor core)
                mbar_arrive(&bars[BAR_FULL + s]);
            }
        }
    } else if (warp == MMA_WARP) {
        // ===================== MMA ISSUER (one thread) =====================
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = BF16, both K-major, N, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            uint32_t it = 0, t = 0;
            for (int64_t st = blockIdx.x; st < n_super; st += gridDim.x, ++t) {
                if (t >= 1) mbar_wait(&bars[BAR_TEMPTY], (t - 1) & 1);            // epilogue drained the accumulators
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&bars[BAR_FULL + s], (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t sA = smem_u32(base + (size_t)s * S::STAGE_BYTES), sB = sA + MT * 3 * A_TILE_BYTES;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t d_main = tmem_base + (uint32_t)(mt * 2 * N), d_corr = d_main + (uint32_t)N;
                        // p = 0: leading product a1*b1 -> main; p = 1..5: corrections (i + j <= 4), smallest first -> corr
                        const int pa[6] = {0, 2, 1, 0, 1, 0}, pb[6] = {0, 0, 1, 2, 0, 1};
#pragma unroll
                        for (int p = 0; p < 6; ++p) {
#pragma unroll
                            for (int ks = 0; ks < KC / 16; ++ks) {
                                const uint64_t da = make_desc(sA + (mt * 3 + pa[p]) * A_TILE_BYTES + ks * 32);
                                const uint64_t db = make_desc(sB + pb[p] * S::B_TILE_BYTES + ks * 32);
                                const bool first = kc == 0 && ks == 0 && (p == 0 || p == 1);
                                umma_bf16(p == 0 ? d_main : d_corr, da, db, idesc, first ? 0u : 1u);
                            }
                        }
                    }
                    umma_commit(&bars[BAR_EMPTY + s]);                 // stage free when these MMAs retire
                }
                umma_commit(&bars[BAR_TFULL]);                         // accumulators complete
            }
        }
    } else {
        // ===================== EPILOGUE: warp w <-> TMEM lanes 32w .. 32w+31 =====================
        uint32_t t = 0;
        for (int64_t st = blockIdx.x; st < n_super; st += gridDim.x, ++t) {
            const int64_t row0 = st * SUPER_M;
            mbar_wait(&bars[BAR_TFULL], t & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int64_t row = row0 + mt * TILE_M + warp * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * 2 * N);
                float as_acc = 0.f, ad_acc = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < N; c0 += 32) {
                    uint32_t v[32], u[32];
#define BGNN_TMEM_LD32(arr, addr)                                                                                              \
                    asm volatile(                                                                                              \
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"            \
                        : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]), \
                          "=r"(arr[8]), "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]), \
                          "=r"(arr[15]), "=r"(arr[16]), "=r"(arr[17]), "=r"(arr[18]), "=r"(arr[19]), "=r"(arr[20]), "=r"(arr[21]), \
                          "=r"(arr[22]), "=r"(arr[23]), "=r"(arr[24]), "=r"(arr[25]), "=r"(arr[26]), "=r"(arr[27]), "=r"(arr[28]), \
                          "=r"(arr[29]), "=r"(arr[30]), "=r"(arr[31])                                                          \
                        : "r"(addr))
                    BGNN_TMEM_LD32(v, taddr + (uint32_t)c0);
                    BGNN_TMEM_LD32(u, taddr + (uint32_t)(N + c0));
#undef BGNN_TMEM_LD32
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                    float x[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __uint_as_float(u[j]);     // main + corrections (RN)
                    if (alpha) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            as_acc = fmaf(x[j], s_att[c0 + j], as_acc);
                            ad_acc = fmaf(x[j], s_att[N + c0 + j], ad_acc);
                        }
                        if (((c0 + 32) % CH) == 0) {          // a head's channels are complete
                            if (row < M) {
                                const int h = (c0 + 32) / CH - 1;
                                alpha[row * (2 * Htot) + h] = as_acc;
                                alpha[row * (2 * Htot) + Htot + h] = ad_acc;
                            }
                            as_acc = 0.f; ad_acc = 0.f;
                        }
                    }
                    if (row < M) {
                        float* dst = Cout + row * N + c0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) st_global_v8(dst + 8 * j, x + 8 * j);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(&bars[BAR_TEMPTY]);
        }
    }
    // teardown
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(S::TMEM_COLS) : "memory");
    }
}

template <int N>
int launch_t(const float* A, int64_t M, int K, const uint16_t* Wsplit, float* C, const float* att_src, const float* att_dst,
             float* alpha, int Htot, int CH, int sms, cudaStream_t s) {
    using S = UmmaSmem<N>;
    static bool configured = false;
    auto kern = proj_umma_kernel<N>;
    if (!configured) {
        BGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES));
        configured = true;
    }
    const int64_t n_super = (M + S::SUPER_M - 1) / S::SUPER_M;
    const unsigned grid = (unsigned)(n_super < sms ? n_super : sms);
    kern<<<grid, UM_THREADS, S::BYTES, s>>>(A, M, K, Wsplit, C, att_src, att_dst, alpha, Htot, CH);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

uint16_t f32_to_bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);      // NaN
    const uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
}
float bf16_to_f32(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

}  // namespace

bool proj_umma_supported(int N, int K, int CH, bool want_alpha) {
    if (!(N == 32 || N == 64 || N == 128 || N == 256)) return false;
    if (K < KC || K % KC != 0) return false;
    if (want_alpha && (CH % 32 != 0 || N % CH != 0)) return false;
    if (want_alpha && N >= 256 && (N / 2) % CH != 0) return false;      // column halves must hold whole heads
    return true;
}

void proj_umma_split_weights(const float* W, int N, int K, uint16_t* out /*[3][N][K]*/) {
    const size_t n = (size_t)N * K;
    for (size_t i = 0; i < n; ++i) {
        const float w = W[i];
        const uint16_t b1 = f32_to_bf16_rn(w);
        const float r1 = w - bf16_to_f32(b1);
        const uint16_t b2 = f32_to_bf16_rn(r1);
        const float r2 = r1 - bf16_to_f32(b2);
        out[i] = b1; out[n + i] = b2; out[2 * n + i] = f32_to_bf16_rn(r2);
    }
}

int launch_proj_umma(const float* A, int64_t M, int K, const uint16_t* Wsplit, int N, float* C, const float* att_src,
                     const float* att_dst, float* alpha, int Htot, int CH, int sms, cudaStream_t s) {
    if (M == 0) return BGNN_OK;
    if (!alpha) { Htot = 1; CH = N; }
    switch (N) {
        case 32: return launch_t<32>(A, M, K, Wsplit, C, att_src, att_dst, alpha, Htot, CH, sms, s);
        case 64: return launch_t<64>(A, M, K, Wsplit, C, att_src, att_dst, alpha, Htot, CH, sms, s);
        case 128: return launch_t<128>(A, M, K, Wsplit, C, att_src, att_dst, alpha, Htot, CH, sms, s);
        case 256: return launch_t<256>(A, M, K, Wsplit, C, att_src, att_dst, alpha, Htot, CH, sms, s);
    }
    bgnn_set_error("proj_umma: unsupported N=%d", N);
    return BGNN_E_UNSUPPORTED;
}

// ---- test hook (declared in include/bgnn.h): C = A . W^T on the tensor-core path, W given as fp32 on the HOST
extern "C" int bgnn_test_proj_umma(const float* A_dev, int64_t M, int K, const float* W_host, int N, float* C_dev,
                                   const float* att_src_dev, const float* att_dst_dev, float* alpha_dev, int heads,
                                   void* stream_) {
    BGNN_CHECK_ARG(A_dev && W_host && C_dev, "null argument");
    const int CH = heads > 0 ? N / heads : N;
    BGNN_CHECK_ARG(proj_umma_supported(N, K, CH, alpha_dev != nullptr), "shape N=%d K=%d not supported by the tcgen05 path", N, K);
    cudaStream_t s = (cudaStream_t)stream_;
    std::vector<uint16_t> split((size_t)3 * N * K);
    proj_umma_split_weights(W_host, N, K, split.data());
    uint16_t* d = nullptr;
    BGNN_CUDA(cudaMalloc(&d, split.size() * 2));
    BGNN_CUDA(cudaMemcpyAsync(d, split.data(), split.size() * 2, cudaMemcpyHostToDevice, s));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int rc = launch_proj_umma(A_dev, M, K, d, N, C_dev, att_src_dev, att_dst_dev, alpha_dev, heads, CH, sms, s);
    cudaStreamSynchronize(s);
    cudaFree(d);
    return rc;
}
