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
//
// Two fusions ride on the same pipeline (template modes):
//   PM_MLP0   the producers compute the A operand on the fly as relu(x W0^T + b0) -- LocalFeatureExtractor's first
//             Linear + ReLU (models/gnn.py:55-68).  Its second Linear has no activation before GATConv.lin, so the host
//             folds it into the layer-0 projection (W' = W_lin W1, bias' = W_lin b1): the [N,64] MLP activations never
//             touch HBM.
//   EM_HEADS  the epilogue applies bias + ReLU to the concatenated first layers of the three heads and finishes them in
//             registers (second layers, softmax / argmax / sigmoid / action; gnn.py:191-260, 392-405, 436-449): one
//             thread per node, only the final per-node outputs are written.
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

// Exact 3-way bf16 split of four fp32 values by TRUNCATION: a1 = top 16 bits of a, r1 = a - a1 (exact), a2 = top 16
// bits of r1, a3 = r1 - a2 (<= 8 significant bits left, so its top 16 bits hold it exactly): a == a1 + a2 + a3.
// Two integer ANDs + two FADDs per value and one PRMT per packed pair -- the producers are issue-bound.
__device__ __forceinline__ void split4(const float4 v, uint2& s1, uint2& s2, uint2& s3) {
    const float x[4] = {v.x, v.y, v.z, v.w};
    uint32_t h1[4], h2[4], h3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h1[i] = __float_as_uint(x[i]);
        const float r1 = x[i] - __uint_as_float(h1[i] & 0xffff0000u);
        h2[i] = __float_as_uint(r1);
        const float r2 = r1 - __uint_as_float(h2[i] & 0xffff0000u);
        h3[i] = __float_as_uint(r2);
    }
    s1.x = __byte_perm(h1[0], h1[1], 0x7632); s1.y = __byte_perm(h1[2], h1[3], 0x7632);
    s2.x = __byte_perm(h2[0], h2[1], 0x7632); s2.y = __byte_perm(h2[2], h2[3], 0x7632);
    s3.x = __byte_perm(h3[0], h3[1], 0x7632); s3.y = __byte_perm(h3[2], h3[3], 0x7632);
}

constexpr int MLP0_MAXK = 128, MLP0_FP = 8;         // PM_MLP0: K <= 128 hidden units, F <= 8 inputs (padded to 8)
constexpr int HEADS_MAXOUT = 10, HEADS_MAXH2 = 64;  // EM_HEADS: <= 8 classes + confidence + correction, h2 <= 64

template <int N, int PMODE = PM_PLAIN, int EMODE = EM_PROJ>
struct UmmaSmem {
    static constexpr int MT = N >= 128 ? 1 : 2;                  // M tiles per super tile
    static constexpr int SUPER_M = MT * TILE_M;
    static constexpr int B_TILE_BYTES = N * KC * 2;
    static constexpr int STAGE_BYTES = MT * 3 * A_TILE_BYTES + 3 * B_TILE_BYTES;
    static constexpr int ACC_COLS = 2 * MT * N;                  // {main, corr} per M tile
    static constexpr int ACC_STAGES = 2 * ACC_COLS <= 512 ? 2 : 1;   // double-buffered accumulators when TMEM allows:
                                                                  // the epilogue of super tile t overlaps the MMAs of t+1
    static constexpr int TMEM_COLS = ACC_COLS * ACC_STAGES <= 32 ? 32 : ACC_COLS * ACC_STAGES <= 64 ? 64 : ACC_COLS * ACC_STAGES <= 128 ? 128
                                     : ACC_COLS * ACC_STAGES <= 256 ? 256 : 512;
    static constexpr int PAR_E = EMODE == EM_PROJ ? 3 * N : N + HEADS_MAXOUT * HEADS_MAXH2 + 16;   // epilogue parameters
    static constexpr int PAR_P = PMODE == PM_MLP0 ? MLP0_MAXK * MLP0_FP + MLP0_MAXK : 0;              // producer parameters
    static constexpr size_t BYTES = (size_t)STAGES * STAGE_BYTES + (PAR_E + PAR_P) * sizeof(float) + 128;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_global_v8(float* p, const float* x) {          // 256-bit store (PTX ISA 8.8, sm_100)
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]),
                 "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]) : "memory");
}

// main + corrections of 32 accumulator columns of this thread's TMEM lane -> x[32] (fp32 add, RN)
__device__ __forceinline__ void tmem_ld_sum32(uint32_t addr_main, uint32_t addr_corr, float* x) {
    uint32_t v[32], u[32];
#define BGNN_TMEM_LD32(arr, addr)                                                                                              \
    asm volatile(                                                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "       \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                            \
        : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]),     \
          "=r"(arr[8]), "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]),             \
          "=r"(arr[15]), "=r"(arr[16]), "=r"(arr[17]), "=r"(arr[18]), "=r"(arr[19]), "=r"(arr[20]), "=r"(arr[21]),           \
          "=r"(arr[22]), "=r"(arr[23]), "=r"(arr[24]), "=r"(arr[25]), "=r"(arr[26]), "=r"(arr[27]), "=r"(arr[28]),           \
          "=r"(arr[29]), "=r"(arr[30]), "=r"(arr[31])                                                                          \
        : "r"(addr))
    BGNN_TMEM_LD32(v, addr_main);
    BGNN_TMEM_LD32(u, addr_corr);
#undef BGNN_TMEM_LD32
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __uint_as_float(u[j]);
}

// warp roles
constexpr int EPI_WARPS = 4, PROD_WARPS = 8, PROD_GROUPS = 2;
constexpr int MMA_WARP = EPI_WARPS, PROD_WARP0 = EPI_WARPS + 1;
constexpr int PROD_GROUP_THREADS = PROD_WARPS / PROD_GROUPS * 32;                  // 128
static_assert(UM_THREADS == (EPI_WARPS + 1 + PROD_WARPS) * 32, "thread count / role mismatch");
// barrier slots
constexpr int BAR_FULL = 0, BAR_EMPTY = STAGES, BAR_TFULL = 2 * STAGES, BAR_TEMPTY = 2 * STAGES + 2, NUM_BARS = 2 * STAGES + 4;

template <int N, int PMODE, int EMODE>
__global__ void __launch_bounds__(UM_THREADS, 1)
proj_umma_kernel(const __grid_constant__ UmmaArgs P) {
    using S = UmmaSmem<N, PMODE, EMODE>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw;                                                      // stage ring (1024 B aligned)
    float* s_par = reinterpret_cast<float*>(base + (size_t)STAGES * S::STAGE_BYTES);      // epilogue parameters
    float* s_prod = s_par + S::PAR_E;                                                     // producer parameters
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_prod + S::PAR_P);                      // NUM_BARS
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

    constexpr int MT = S::MT, SUPER_M = S::SUPER_M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t M = P.M;
    const int K = P.K;
    const int nkc = K / KC;
    const int64_t n_super = (M + SUPER_M - 1) / SUPER_M;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bars[BAR_FULL + s], PROD_GROUP_THREADS); mbar_init(&bars[BAR_EMPTY + s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bars[BAR_TFULL + a], 1); mbar_init(&bars[BAR_TEMPTY + a], EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (EMODE == EM_PROJ) {          // s_par = att_src [N] | att_dst [N] | bias [N]
        for (int i = tid; i < N; i += UM_THREADS) {
            s_par[i] = P.alpha ? __ldg(P.att_src + i) : 0.f;
            s_par[N + i] = P.alpha ? __ldg(P.att_dst + i) : 0.f;
            s_par[2 * N + i] = P.bias ? __ldg(P.bias + i) : 0.f;
        }
    } else {                         // s_par = hb0 [N, zero padded] | W2 [nout][h2] (cls rows, conf, corr) | b2 [nout]
        const int real = P.nblk * P.h2;
        for (int i = tid; i < N; i += UM_THREADS) s_par[i] = i < real ? __ldg(P.hb0 + i) : 0.f;
        float* w2 = s_par + N;
        float* b2 = w2 + HEADS_MAXOUT * HEADS_MAXH2;
        for (int i = tid; i < (P.ncls + 2) * P.h2; i += UM_THREADS) {
            const int q = i / P.h2, u = i - q * P.h2;
            float v = 0.f;
            if (q < P.ncls) v = __ldg(P.Wcls + q * P.h2 + u);
            else if (q == P.ncls) v = __ldg(P.Wconf + u);
            else if (P.nblk > 2) v = __ldg(P.Wcorr + u);
            w2[i] = v;
        }
        for (int i = tid; i < P.ncls + 2; i += UM_THREADS)
            b2[i] = i < P.ncls ? __ldg(P.bcls + i) : i == P.ncls ? __ldg(P.bconf) : (P.nblk > 2 ? __ldg(P.bcorr) : 0.f);
    }
    if (PMODE == PM_MLP0) {          // s_prod = W0 [K][8] (F padded with zeros) | b0 [K]
        for (int i = tid; i < K * MLP0_FP; i += UM_THREADS) {
            const int k = i / MLP0_FP, f = i % MLP0_FP;
            s_prod[i] = f < P.F ? __ldg(P.W0 + k * P.F + f) : 0.f;
        }
        for (int i = tid; i < K; i += UM_THREADS) s_prod[MLP0_MAXK * MLP0_FP + i] = __ldg(P.b0 + i);
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
        // ===================== PRODUCERS: two groups of 128 threads own alternate K chunks =====================
        const int grp = (warp - PROD_WARP0) / (PROD_WARPS / PROD_GROUPS);
        const int ptid = tid - (PROD_WARP0 + grp * (PROD_WARPS / PROD_GROUPS)) * 32;      // 0..127
        constexpr int A_F4 = MT * TILE_M * (KC / 4) / PROD_GROUP_THREADS;                  // 8 or 16 float4 per thread and chunk
        uint32_t it = 0;
        for (int64_t st = blockIdx.x; st < n_super; st += gridDim.x) {
            const int64_t row0 = st * SUPER_M;
            for (int kc = 0; kc < nkc; ++kc, ++it) {
                if ((int)(it % PROD_GROUPS) != grp) continue;
                const int s = it % STAGES;
                unsigned char* stage = base + (size_t)s * S::STAGE_BYTES;
                constexpr uint32_t W_BYTES = 3 * S::B_TILE_BYTES;
                auto issue_w = [&]() {  // W chunk: the host pre-arranged [kc][split] tiles in the swizzled smem layout -> ONE bulk copy
                    const uint32_t bar = smem_u32(&bars[BAR_FULL + s]);
                    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(W_BYTES) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                                 ::"r"(smem_u32(stage + MT * 3 * A_TILE_BYTES)),
                                   "l"(reinterpret_cast<const unsigned char*>(P.Wsplit) + (size_t)kc * W_BYTES), "r"(W_BYTES), "r"(bar)
                                 : "memory");
                };
                if (PMODE == PM_PLAIN) {
                    // A values first (long latency, registers only), then wait for the stage to be free
                    float4 av[A_F4];
#pragma unroll
                    for (int i = 0; i < A_F4; ++i) {
                        const int f = ptid + i * PROD_GROUP_THREADS;
                        const int mt = f >> 10, rem = f & 1023, r = rem >> 3, c4 = rem & 7;
                        const int64_t gr = row0 + mt * TILE_M + r;
                        av[i] = gr < M ? __ldg(reinterpret_cast<const float4*>(P.A + gr * K + kc * KC + c4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (it >= STAGES) mbar_wait(&bars[BAR_EMPTY + s], ((it / STAGES) - 1) & 1);
                    if (ptid == 0) issue_w();
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
                } else {
                    // thread <-> node: its F inputs are loaded once, then the 32 hidden units of this chunk are
                    // relu(b0 + x W0^T) straight into the split tiles (W0 rows are warp-uniform broadcasts from smem)
                    float xr[MT][MLP0_FP];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const int64_t gr = row0 + mt * TILE_M + ptid;
#pragma unroll
                        for (int q = 0; q < MLP0_FP; ++q) xr[mt][q] = (gr < M && q < P.F) ? __ldg(P.x + gr * P.F + q) : 0.f;
                    }
                    if (it >= STAGES) mbar_wait(&bars[BAR_EMPTY + s], ((it / STAGES) - 1) & 1);
                    if (ptid == 0) issue_w();
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll 2
                        for (int c4 = 0; c4 < KC / 4; ++c4) {
                            const int k0 = kc * KC + c4 * 4;
                            float o[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float4 w0 = *reinterpret_cast<const float4*>(s_prod + (k0 + u) * MLP0_FP);
                                const float4 w1 = *reinterpret_cast<const float4*>(s_prod + (k0 + u) * MLP0_FP + 4);
                                float a = xr[mt][0] * w0.x;
                                a = fmaf(xr[mt][1], w0.y, a); a = fmaf(xr[mt][2], w0.z, a); a = fmaf(xr[mt][3], w0.w, a);
                                a = fmaf(xr[mt][4], w1.x, a); a = fmaf(xr[mt][5], w1.y, a); a = fmaf(xr[mt][6], w1.z, a);
                                a = fmaf(xr[mt][7], w1.w, a);
                                o[u] = fmaxf(a + s_prod[MLP0_MAXK * MLP0_FP + k0 + u], 0.f);
                            }
                            uint2 s1, s2, s3;
                            split4(make_float4(o[0], o[1], o[2], o[3]), s1, s2, s3);
                            unsigned char* t0 = stage + (size_t)(mt * 3) * A_TILE_BYTES + sw64_off(ptid, c4 >> 1) + (c4 & 1) * 8;
                            *reinterpret_cast<uint2*>(t0) = s1;
                            *reinterpret_cast<uint2*>(t0 + A_TILE_BYTES) = s2;
                            *reinterpret_cast<uint2*>(t0 + 2 * A_TILE_BYTES) = s3;
                        }
                    }
                }
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
                const uint32_t as = t % S::ACC_STAGES;
                if (t >= (uint32_t)S::ACC_STAGES) mbar_wait(&bars[BAR_TEMPTY + as], (t / S::ACC_STAGES - 1) & 1);   // epilogue drained this stage
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&bars[BAR_FULL + s], (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t sA = smem_u32(base + (size_t)s * S::STAGE_BYTES), sB = sA + MT * 3 * A_TILE_BYTES;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t d_main = tmem_base + as * S::ACC_COLS + (uint32_t)(mt * 2 * N), d_corr = d_main + (uint32_t)N;
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
                umma_commit(&bars[BAR_TFULL + as]);                    // accumulators complete
            }
        }
    } else {
        // ===================== EPILOGUE: warp w <-> TMEM lanes 32w .. 32w+31, one thread per output row =====================
        uint32_t t = 0;
        for (int64_t st = blockIdx.x; st < n_super; st += gridDim.x, ++t) {
            const int64_t row0 = st * SUPER_M;
            const uint32_t as = t % S::ACC_STAGES;
            mbar_wait(&bars[BAR_TFULL + as], (t / S::ACC_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int64_t row = row0 + mt * TILE_M + warp * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * S::ACC_COLS + (uint32_t)(mt * 2 * N);
                if (EMODE == EM_PROJ) {
                    float as_acc = 0.f, ad_acc = 0.f;
#pragma unroll 1
                    for (int c0 = 0; c0 < N; c0 += 32) {
                        float x[32];
                        tmem_ld_sum32(taddr + (uint32_t)c0, taddr + (uint32_t)(N + c0), x);
                        if (P.bias) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) x[j] += s_par[2 * N + c0 + j];
                        }
                        if (P.alpha) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                as_acc = fmaf(x[j], s_par[c0 + j], as_acc);
                                ad_acc = fmaf(x[j], s_par[N + c0 + j], ad_acc);
                            }
                            if (((c0 + 32) % P.CH) == 0) {          // a head's channels are complete
                                if (row < M) {
                                    const int h = (c0 + 32) / P.CH - 1;
                                    P.alpha[row * (2 * P.Htot) + h] = as_acc;
                                    P.alpha[row * (2 * P.Htot) + P.Htot + h] = ad_acc;
                                }
                                as_acc = 0.f; ad_acc = 0.f;
                            }
                        }
                        if (row < M) {
                            float* dst = P.Cout + row * N + c0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) st_global_v8(dst + 8 * j, x + 8 * j);
                        }
                    }
                } else {
                    // heads: columns [b*h2, (b+1)*h2) = hidden units of head b (0 cls, 1 conf, 2 corr), h2 % 32 == 0
                    const float* w2 = s_par + N;
                    const float* b2 = w2 + HEADS_MAXOUT * HEADS_MAXH2;
                    const int h2 = P.h2, ncls = P.ncls, real = P.nblk * h2;
                    float lg[8], pc = 0.f, pr = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) lg[k] = 0.f;
#pragma unroll 1
                    for (int c0 = 0; c0 < real; c0 += 32) {
                        float x[32];
                        tmem_ld_sum32(taddr + (uint32_t)c0, taddr + (uint32_t)(N + c0), x);
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j] + s_par[c0 + j], 0.f);
                        const int blk = c0 / h2, u0 = c0 - blk * h2;
                        if (blk == 0) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (k < ncls) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) lg[k] = fmaf(x[j], w2[k * h2 + u0 + j], lg[k]);
                                }
                        } else if (blk == 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) pc = fmaf(x[j], w2[ncls * h2 + u0 + j], pc);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) pr = fmaf(x[j], w2[(ncls + 1) * h2 + u0 + j], pr);
                        }
                    }
                    if (row < M) {
                        float best = -CUDART_INF_F;
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < ncls) { lg[k] += b2[k]; best = fmaxf(best, lg[k]); }
                        float pb[8], sum = 0.f;
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < ncls) { pb[k] = expf(lg[k] - best); sum += pb[k]; }
                        int arg = 0; float pmax = -1.f;
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < ncls) {
                                pb[k] = pb[k] / sum;
                                if (pb[k] > pmax) { pmax = pb[k]; arg = k; }      // first max wins (torch.argmax)
                                if (P.logits) P.logits[row * ncls + k] = lg[k];
                                if (P.probs) P.probs[row * ncls + k] = pb[k];
                            }
                        const float cf = 1.f / (1.f + expf(-(pc + b2[ncls])));
                        if (P.cls) P.cls[row] = arg;
                        if (P.conf) P.conf[row] = cf;
                        if (P.corr && P.nblk > 2) P.corr[row] = pr + b2[ncls + 1];
                        if (P.action) {
                            int64_t a = (arg == 2 && cf > P.thr_auto) ? 1 : 0;          // gnn.py:439-441
                            if (cf < P.thr_review) a = 2;                                // gnn.py:444-445
                            P.action[row] = a;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(&bars[BAR_TEMPTY + as]);
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

template <int N, int PMODE, int EMODE>
int launch_t(const UmmaArgs& a, int sms, cudaStream_t s) {
    using S = UmmaSmem<N, PMODE, EMODE>;
    static_assert(S::BYTES <= 232448, "shared memory budget exceeded");
    static bool configured[64] = {};                    // the attribute is per device
    auto kern = proj_umma_kernel<N, PMODE, EMODE>;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        BGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES));
        configured[dev & 63] = true;
    }
    const int64_t n_super = (a.M + S::SUPER_M - 1) / S::SUPER_M;
    const unsigned grid = (unsigned)(n_super < sms ? n_super : sms);
    kern<<<grid, UM_THREADS, S::BYTES, s>>>(a);
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

// host: W [N,K] fp32 -> bf16 splits w1 + w2 + w3 (round to nearest), arranged as [K/32][3][N x 32 tile] where each tile
// is ALREADY in the K-major SWIZZLE_64B shared-memory layout the MMA descriptors expect (one bulk copy per K chunk).
static size_t sw64_off_host(int r, int c) { return (size_t)(r >> 3) * SBO_BYTES + (r & 7) * 64 + ((c ^ ((r & 7) >> 1)) << 4); }
void proj_umma_split_weights(const float* W, int N, int K, uint16_t* out /*[K/32][3][N*32]*/) {
    const size_t tile = (size_t)N * KC;
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const float w = W[(size_t)n * K + k];
            const uint16_t b1 = f32_to_bf16_rn(w);
            const float r1 = w - bf16_to_f32(b1);
            const uint16_t b2 = f32_to_bf16_rn(r1);
            const float r2 = r1 - bf16_to_f32(b2);
            const uint16_t b3 = f32_to_bf16_rn(r2);
            const int kc = k / KC, kk = k % KC;
            const size_t e = sw64_off_host(n, kk >> 3) / 2 + (kk & 7);
            uint16_t* base = out + (size_t)kc * 3 * tile;
            base[e] = b1; base[tile + e] = b2; base[2 * tile + e] = b3;
        }
}

bool proj_umma_mlp0_supported(int N, int K, int F) {
    return (N == 128 || N == 256) && K % KC == 0 && K >= KC && K <= MLP0_MAXK && F >= 1 && F <= MLP0_FP;
}

bool proj_umma_heads_supported(int K, int h2, int nblk, int ncls, int* Npad) {
    const int real = nblk * h2;
    const int np = real <= 128 ? 128 : 256;
    if (Npad) *Npad = np;
    return h2 % 32 == 0 && h2 <= HEADS_MAXH2 && real <= 256 && ncls >= 1 && ncls <= 8 && K % KC == 0 && K >= KC;
}

int launch_umma(const UmmaArgs& a, int N, int pmode, int emode, int sms, cudaStream_t s) {
    if (a.M == 0) return BGNN_OK;
    if (pmode == PM_PLAIN && emode == EM_PROJ) {
        switch (N) {
            case 32: return launch_t<32, PM_PLAIN, EM_PROJ>(a, sms, s);
            case 64: return launch_t<64, PM_PLAIN, EM_PROJ>(a, sms, s);
            case 128: return launch_t<128, PM_PLAIN, EM_PROJ>(a, sms, s);
            case 256: return launch_t<256, PM_PLAIN, EM_PROJ>(a, sms, s);
        }
    } else if (pmode == PM_MLP0 && emode == EM_PROJ) {
        switch (N) {
            case 128: return launch_t<128, PM_MLP0, EM_PROJ>(a, sms, s);
            case 256: return launch_t<256, PM_MLP0, EM_PROJ>(a, sms, s);
        }
    } else if (pmode == PM_PLAIN && emode == EM_HEADS) {
        switch (N) {
            case 128: return launch_t<128, PM_PLAIN, EM_HEADS>(a, sms, s);
            case 256: return launch_t<256, PM_PLAIN, EM_HEADS>(a, sms, s);
        }
    }
    bgnn_set_error("proj_umma: unsupported N=%d (producer mode %d, epilogue mode %d)", N, pmode, emode);
    return BGNN_E_UNSUPPORTED;
}

int launch_proj_umma(const float* A, int64_t M, int K, const uint16_t* Wsplit, int N, float* C, const float* att_src,
                     const float* att_dst, float* alpha, int Htot, int CH, int sms, cudaStream_t s) {
    UmmaArgs a = {};
    a.A = A; a.M = M; a.K = K; a.Wsplit = Wsplit; a.Cout = C; a.att_src = att_src; a.att_dst = att_dst; a.alpha = alpha;
    a.Htot = alpha ? Htot : 1; a.CH = alpha ? CH : N;
    return launch_umma(a, N, PM_PLAIN, EM_PROJ, sms, s);
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
