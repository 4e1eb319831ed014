// Projection GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//     C[M, N] = A[M, K] . W[N, K]^T      (fp32 in, fp32 out)     + fused epilogues
//
// This is the `h = x W^T` of every GATConv layer (PyG GATConv.lin, called from models/gnn.py:175-176) and the first
// layers of the output heads (gnn.py:191-260): skinny dense contractions, M = nodes (millions), N, K <= 256.
//
// Precision.  fp32 parity (atol 2e-5 on the logits) rules out plain fp16 / bf16 / tf32 operands.  Both operands are
// split into TWO fp16 terms (Ootomo & Yokota 2022, "halfhalf"):  a*s = a_hi + a_lo * 2^-11 with a_hi = fp16(a*s),
// a_lo = fp16((a*s - a_hi) * 2^11), 22 significant bits in total, and the three products a_hi*b_hi, a_hi*b_lo, a_lo*b_hi
// are accumulated in fp32 in TMEM -- half the tensor work of a 3-term bf16 split at the same fp32-grade accuracy
// (dropped terms ~2^-22).  The tensor core accumulates with truncation, so the leading product gets its OWN accumulator
// (K/16 chained MMAs) and the two correction products a second one; the epilogue adds them in fp32 (RN) as
// main + corr * 2^-11.  fp16 has 5 exponent bits: each operand is scaled by a power of two `s` so that |a*s| < 1.  For W
// the host knows the maximum; for A the kernel derives it from a device-side UPPER BOUND of max|A| that the previous
// kernel left behind (every projection epilogue tracks max|C|; the layer in between is a convex combination followed by
// an affine map, so |A| <= bn_mul * max|C| + bn_add).  Scaling by powers of two is exact and is undone in the epilogue.
//
// One persistent CTA per SM, 608 threads, warp-specialised:
//   A loader (1 thread): each fp32 A chunk [MT*128 x 32] is fetched by ONE TMA tensor copy (cp.async.bulk.tensor.2d,
//       mbarrier complete_tx) into a raw ring 2-6 chunks ahead: HBM latency is hidden by the async engine, not by registers.
//   W loader (1 thread): the W chunk (pre-split and pre-swizzled on the host) arrives by ONE cp.async.bulk per chunk in
//       its own 3-deep ring (its ~1000-cycle L2 latency must not sit between "stage free" and "stage full").
//   producers (8 warps, two groups owning alternate K chunks): raw chunk -> 2-way fp16 split -> STS into K-major
//       SWIZZLE_64B tiles.
//   MMA issue (one thread): 6 x tcgen05.mma.kind::f16 (M=128, N, K=16) per M tile and chunk, tcgen05.commit -> mbarriers
//       free the A stage and the W slot.  For N = K = 256 the 48 MMAs of a tile are 6 500 cycles of tensor time and the
//       epilogue another 6 200: main + correction accumulators fill all 512 TMEM columns, so the two cannot overlap.
//   epilogue (8 warps; warp w <-> TMEM lanes 32*(w%4).., column half w/4 when N = 256): tcgen05.ld 32x32b.x32 of main
//       and corr, rescale, bias, alpha_src / alpha_dst dot products per head; each [32 x 32] chunk is staged in a
//       SWIZZLE_128B shared-memory tile and written by ONE TMA tensor store (thread-per-row global stores touch 32
//       different lines per instruction and made the kernel LSU-bound).  For N <= 128 the accumulators are
//       double-buffered so the epilogue of one tile overlaps the MMAs of the next.
//
// Two fusions ride on the same pipeline (template modes):
//   PM_MLP0   the producers compute the A operand on the fly as relu(x W0^T + b0) -- LocalFeatureExtractor's first
//             Linear + ReLU (models/gnn.py:55-68).  Its second Linear has no activation before GATConv.lin, so the host
//             folds it into the layer-0 projection (W' = W_lin W1, bias' = W_lin b1): the [N,64] MLP activations never
//             touch HBM.
//   EM_HEADS  the epilogue applies bias + ReLU to the concatenated first layers of the three heads and finishes them in
//             registers (second layers, softmax / argmax / sigmoid / action; gnn.py:191-260, 392-405, 436-449): one
//             thread per node, only the final per-node outputs are written.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "proj_umma.cuh"

#ifdef BGNN_AGG_TIMING
__device__ unsigned long long g_umma_t[12];
#define UT_DECL() long long ut_prev = clock64()
#ifndef BGNN_UT_FILTER
#define BGNN_UT_FILTER true            // e.g. -DBGNN_UT_FILTER="(PMODE==PM_MLP0)": count only that instantiation
#endif
#define UT_TICK(slot) do { long long ut_now = clock64(); if (BGNN_UT_FILTER) atomicAdd(&g_umma_t[slot], (unsigned long long)(ut_now - ut_prev)); ut_prev = ut_now; } while (0)
extern "C" int bgnn_debug_umma_timing(unsigned long long* out8, int reset) {
    if (out8) cudaMemcpyFromSymbol(out8, g_umma_t, sizeof(unsigned long long) * 12);
    if (reset) { unsigned long long z[12] = {}; cudaMemcpyToSymbol(g_umma_t, z, sizeof(z)); }
    return 0;
}
#define UT_TICKP(slot) do { long long ut_now = clock64(); atomicAdd(&g_umma_t[slot], (unsigned long long)(ut_now - ut_prev)); ut_prev = ut_now; } while (0)
#else
#define UT_DECL()
#define UT_TICK(slot)
#define UT_TICKP(slot)
#endif

namespace {

constexpr int KC = 32;                 // K elements per stage (64 B of fp16 = one SWIZZLE_64B row)
constexpr int STAGES = 6;              // fp16 operand stages, upper bound (barrier slots); per-shape depth in UmmaSmem
constexpr int TILE_M = 128;
constexpr int NSPLIT = 2;              // hi, lo
constexpr int A_TILE_BYTES = TILE_M * KC * 2;     // 8 KB per split
constexpr uint32_t SBO_BYTES = 8 * KC * 2;        // 512 B between 8-row groups
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

// warp roles
constexpr int EPI_WARPS = 8, PROD_WARPS = 8, PROD_GROUPS = 2;
constexpr int MMA_WARP = EPI_WARPS, LOAD_WARP = EPI_WARPS + 1, WLOAD_WARP = EPI_WARPS + 2, PROD_WARP0 = EPI_WARPS + 3;
constexpr int UM_THREADS = (EPI_WARPS + 3 + PROD_WARPS) * 32;                      // 608
constexpr int PROD_GROUP_THREADS = PROD_WARPS / PROD_GROUPS * 32;                  // 128
constexpr int MAX_RAW = 6;             // raw fp32 A ring (bulk-copied rows): depth chosen per shape
// barrier slots
constexpr int NW = 3;                  // W chunk ring (own, deeper ring: the 32 KB copy has ~1000 cycles of L2 latency)
constexpr int BAR_FULL = 0, BAR_EMPTY = STAGES, BAR_TFULL = 2 * STAGES, BAR_TEMPTY = 2 * STAGES + 2, BAR_RFULL = 2 * STAGES + 4,
              BAR_REMPTY = BAR_RFULL + MAX_RAW, BAR_WFULL = BAR_REMPTY + MAX_RAW, BAR_WEMPTY = BAR_WFULL + NW,
              NUM_BARS = BAR_WEMPTY + NW;

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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at the same shared-memory offset in every CTA of `mask` when the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of the 16 B chunk holding elements [8c, 8c+8) of row r inside a K-major SWIZZLE_64B tile
__device__ __forceinline__ uint32_t sw64_off(int r, int c) {
    return (uint32_t)((r >> 3) * SBO_BYTES + (r & 7) * 64 + ((c ^ ((r & 7) >> 1)) << 4));
}

// 2-way fp16 split of four scaled fp32 values: hi = fp16(a*s), lo = fp16((a*s - hi) * 2^11)  (a*s - hi is exact)
__device__ __forceinline__ void split4h(const float4 v, float s, uint2& hi, uint2& lo) {
    const float a0 = v.x * s, a1 = v.y * s, a2 = v.z * s, a3 = v.w * s;
    const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn((a0 - f01.x) * LO_SCALE, (a1 - f01.y) * LO_SCALE);
    const __half2 l23 = __floats2half2_rn((a2 - f23.x) * LO_SCALE, (a3 - f23.y) * LO_SCALE);
    hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
    lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}

constexpr int MLP0_MAXK = 128, MLP0_FP = 8;         // PM_MLP0: K <= 128 hidden units, F <= 8 inputs (padded to 8)
constexpr int HEADS_MAXOUT = 10, HEADS_MAXH2 = 64;  // EM_HEADS: <= 8 classes + confidence + correction, h2 <= 64

// NH = 2 ("paired CTAs, resident W"): the output is NT = 2 N columns wide; CTA b owns column block b & 1 for the whole launch and
// walks the M tiles (b >> 1) + k * gridDim.x / 2, so the two CTAs of a pair work on the same A tile at about the same time
// (the second fetch is an L2 hit).  The CTA's half of the split weights (K <= 256: 8 chunks x 16 KB) is loaded ONCE and
// stays in shared memory -- with a 128 x 256 tile per CTA the 256 KB of W re-read from L2 for every 128 rows were twice
// the A traffic and put the kernel against the L2 throughput cap -- and its accumulators fit TMEM twice, so the epilogue
// of tile t overlaps the MMAs of tile t + 1.
template <int N, int PMODE = PM_PLAIN, int EMODE = EM_PROJ, int NH = 1>
struct UmmaSmem {
    static constexpr int NT = N * NH;                            // total output width
    static constexpr int MT = N >= 128 ? 1 : 2;                  // M tiles per super tile
    static constexpr int SUPER_M = MT * TILE_M;
    static constexpr int B_TILE_BYTES = N * KC * 2;
    static constexpr int STAGE_BYTES = MT * NSPLIT * A_TILE_BYTES;       // A operand stage (hi, lo tiles per M tile)
    static constexpr int W_SLOT_BYTES = NSPLIT * B_TILE_BYTES;           // W chunk (hi, lo tiles)
    static constexpr int ACC_COLS = 2 * MT * N;                  // {main, corr} per M tile
    static constexpr int ACC_STAGES = 2 * ACC_COLS <= 512 ? 2 : 1;   // double-buffered accumulators when TMEM allows:
                                                                  // the epilogue of super tile t overlaps the MMAs of t+1
    static constexpr int TMEM_COLS = ACC_COLS * ACC_STAGES <= 32 ? 32 : ACC_COLS * ACC_STAGES <= 64 ? 64 : ACC_COLS * ACC_STAGES <= 128 ? 128
                                     : ACC_COLS * ACC_STAGES <= 256 ? 256 : 512;
    // epilogue warps 4..7 share the TMEM lanes of warps 0..3 and take the other half of the work: the upper column half
    // (EM_PROJ, N = 256) or the correction head (EM_HEADS; one warp per scheduler cannot hide its own FMA latency)
    static constexpr int EPI_SPLIT = ((EMODE == EM_PROJ && NT == 256) || EMODE == EM_HEADS) ? 2 : 1;
    // EM_PROJ: every epilogue warp stages its [32 rows x 32 cols] chunk in a SWIZZLE_128B tile and hands it to a TMA
    // tensor store (one 4 KB store instead of 128 scattered 32 B sectors through the LSU)
#ifndef BGNN_N256_TMA_STORE
#define BGNN_N256_TMA_STORE 0
#endif
    // N = 256 from global A: shared memory holds either the staging tiles (with a 2-deep raw ring) or a 4-deep raw ring
#ifndef BGNN_SPLIT_TMA_STORE
#define BGNN_SPLIT_TMA_STORE 1
#endif
    // (PM_SPLIT, measured: the paired N = 2 x 128 variant prefers direct stores and a 6-deep operand ring, 1.71 vs 1.82 ms;
    //  N = 64 prefers the TMA stores, 0.79 vs 0.85 ms)
    static constexpr bool TMA_STORE = EMODE == EM_PROJ && (PMODE == PM_SPLIT ? (BGNN_SPLIT_TMA_STORE != 0 && NH == 1)
                                                          : !(PMODE == PM_PLAIN && N == 256 && !BGNN_N256_TMA_STORE) && (NH == 1 || PMODE == PM_MLP0));
    static constexpr int STG_WARPS = TMA_STORE ? 4 * EPI_SPLIT : 0;
    static constexpr int STG_BYTES = STG_WARPS * 4096;
    static constexpr int NSTAGE = PMODE == PM_SPLIT ? (TMA_STORE ? 4 : 6) : ((PMODE == PM_PLAIN && N == 256) || NH > 1) ? 2 : 3;     // fp16 operand ring depth (measured: for
                                                                                     // N = 256 a deeper raw ring beats a third stage)
    static constexpr int PAR_E = EMODE == EM_PROJ ? 3 * (NH > 1 ? N : NT) : N + HEADS_MAXOUT * HEADS_MAXH2 + 16 + 2 * TILE_M;  // epilogue parameters (+ class hand-over)
    static constexpr int PAR_P = PMODE == PM_MLP0 ? MLP0_MAXK * MLP0_FP + MLP0_MAXK : 0;              // producer parameters
    // PM_PLAIN: the A chunk [SUPER_M x 32] fp32 is fetched by TMA into a raw ring, RAW_STAGES chunks ahead
    static constexpr int RAW_BYTES = PMODE == PM_PLAIN ? SUPER_M * KC * 4 : 0;
    // The depth must be a multiple of the producer-group count so that a raw slot is always consumed by the SAME group:
    // TMA copies may complete out of order, and a group that last saw the slot two phases ago would otherwise pass the
    // parity wait of a phase that has not completed (then over-arrive on the empty barrier -> launch failure).
    static constexpr int RAW_STAGES = PMODE != PM_PLAIN ? 0 : (N == 256 ? (BGNN_N256_TMA_STORE ? 2 : 4) : N == 128 ? (NH > 1 ? 4 : 6) : 2);
    static constexpr int W_SLOTS = NH > 1 ? (PMODE == PM_MLP0 ? MLP0_MAXK / KC : 8) : NW;   // NH = 2: the whole K extent stays resident
    static constexpr int PARW = NH > 1 ? N : NT;                          // epilogue parameter columns held by this CTA
    static_assert(NSTAGE <= STAGES, "barrier slots");
    static_assert(RAW_STAGES % PROD_GROUPS == 0, "raw ring depth must be a multiple of the producer group count");
    static constexpr size_t BYTES = (size_t)NSTAGE * STAGE_BYTES + (size_t)W_SLOTS * W_SLOT_BYTES + (size_t)RAW_STAGES * RAW_BYTES +
                                    STG_BYTES + (PAR_E + PAR_P) * sizeof(float) + NUM_BARS * 8 + 16;
};

#ifndef BGNN_EPI_PACKED
#define BGNN_EPI_PACKED 1
#endif

__device__ __forceinline__ void st_global_v8(float* p, const float* x) {          // 256-bit store (PTX ISA 8.8, sm_100)
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]),
                 "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]) : "memory");
}

// packed fp32 pairs (sm_100 fma/add/mul.rn.f32x2 = two IEEE operations per issue slot; the epilogue is one long dependent
// instruction stream per warp, so halving its length shortens the critical path)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void unpk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// (main + corr * 2^-11) * scale of 32 accumulator columns of this thread's TMEM lane -> x[32]
__device__ __forceinline__ void tmem_ld_sum32(uint32_t addr_main, uint32_t addr_corr, float scale, float* x) {
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
#if BGNN_EPI_PACKED
    const uint64_t inv2 = pk2(LO_INV, LO_INV), sc2 = pk2(scale, scale);
#pragma unroll
    for (int j = 0; j < 32; j += 2) unpk2(fmul2(ffma2(pk2u(u[j], u[j + 1]), inv2, pk2u(v[j], v[j + 1])), sc2), x[j], x[j + 1]);
#else
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = fmaf(__uint_as_float(u[j]), LO_INV, __uint_as_float(v[j])) * scale;
#endif
}

template <int N, int PMODE, int EMODE, int NH>
__global__ void __launch_bounds__(UM_THREADS, 1)
proj_umma_kernel(const __grid_constant__ UmmaArgs P, const __grid_constant__ CUtensorMap tmapA,
                 const __grid_constant__ CUtensorMap tmapA2, const __grid_constant__ CUtensorMap tmapC) {
    using S = UmmaSmem<N, PMODE, EMODE, NH>;
    constexpr int NT = S::NT;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw;                                                      // stage ring (1024 B aligned)
    unsigned char* wring = base + (size_t)S::NSTAGE * S::STAGE_BYTES;                    // W chunk ring
    unsigned char* raw = wring + (size_t)S::W_SLOTS * S::W_SLOT_BYTES;                           // raw fp32 A ring
    unsigned char* stg = raw + (size_t)S::RAW_STAGES * S::RAW_BYTES;                     // epilogue store staging (1024 B aligned)
    float* s_par = reinterpret_cast<float*>(stg + S::STG_BYTES);                          // epilogue parameters
    float* s_prod = s_par + S::PAR_E;                                                     // producer parameters
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_prod + S::PAR_P);                      // NUM_BARS
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

    constexpr int MT = S::MT, SUPER_M = S::SUPER_M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t M = P.M;
    const int K = P.K;
    const int nkc = K / KC;
    const int64_t n_super = (M + SUPER_M - 1) / SUPER_M;
    // tile walk of this CTA (NH = 2: pairs of CTAs share the M tiles, each CTA keeps one column block)
    const int nh = NH > 1 ? (int)(blockIdx.x & 1) : 0;
    const int64_t st_first = NH > 1 ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
    const int64_t st_step = NH > 1 ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
    const int cg0 = nh * N;                                    // first global output column of this CTA
    // 2-CTA cluster with multicast A (PM_SPLIT pairs): a stage is free when BOTH CTAs' MMAs have retired (empty count 2)
    const bool mc = (PMODE == PM_SPLIT && NH == 2) && P.mc != 0;

    // power-of-two operand scale from the upper bound of max|A|: |a * sA| < 1 (exact; undone in the epilogue)
    float sA = 1.f, out_scale = P.w_inv_scale;
    {
        float bound = P.amax_add;
        if (P.amax_src) bound = fmaf(P.amax_mul, __ldg(P.amax_src), P.amax_add);
        const int e = (int)((__float_as_uint(bound) >> 23) & 0xffu);
        const int es = 253 - e;                                     // biased exponent of 2^-(floor(log2 bound) + 1)
        if (e > 0 && e < 255 && es >= 1 && es <= 253) {
            sA = __uint_as_float((uint32_t)es << 23);
            out_scale *= __uint_as_float((uint32_t)(254 - es) << 23);
        }
    }

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bars[BAR_FULL + s], PMODE == PM_SPLIT ? 1 : PROD_GROUP_THREADS); mbar_init(&bars[BAR_EMPTY + s], mc ? 2 : 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bars[BAR_TFULL + a], 1); mbar_init(&bars[BAR_TEMPTY + a], 4 * 32 * S::EPI_SPLIT); }
        for (int r = 0; r < MAX_RAW; ++r) { mbar_init(&bars[BAR_RFULL + r], 1); mbar_init(&bars[BAR_REMPTY + r], PROD_GROUP_THREADS); }
        for (int w = 0; w < NW; ++w) { mbar_init(&bars[BAR_WFULL + w], 1); mbar_init(&bars[BAR_WEMPTY + w], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (EMODE == EM_PROJ) {          // s_par = att_src [N] | att_dst [N] | bias [N]
        for (int i = tid; i < S::PARW; i += UM_THREADS) {
            s_par[i] = P.alpha ? __ldg(P.att_src + cg0 + i) : 0.f;
            s_par[S::PARW + i] = P.alpha ? __ldg(P.att_dst + cg0 + i) : 0.f;
            s_par[2 * S::PARW + i] = P.bias ? __ldg(P.bias + cg0 + i) : 0.f;
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
    if (mc) cluster_sync_all();          // the peer's barriers are initialised before anything is multicast to them

    if (warp >= PROD_WARP0) {
        if constexpr (PMODE != PM_SPLIT) {          // (PM_SPLIT: the operand arrives already split, the TMA fills the stages)
        // ===================== PRODUCERS: two groups of 128 threads own alternate K chunks =====================
        const int grp = (warp - PROD_WARP0) / (PROD_WARPS / PROD_GROUPS);
        const int ptid = tid - (PROD_WARP0 + grp * (PROD_WARPS / PROD_GROUPS)) * 32;      // 0..127
        constexpr int A_F4 = MT * TILE_M * (KC / 4) / PROD_GROUP_THREADS;                  // 8 or 16 float4 per thread and chunk
        uint32_t it = 0;
        for (int64_t st = st_first; st < n_super; st += st_step) {
            const int64_t row0 = st * SUPER_M;
            for (int kc = 0; kc < nkc; ++kc, ++it) {
                if ((int)(it % PROD_GROUPS) != grp) continue;
                const int s = it % S::NSTAGE;
                unsigned char* stage = base + (size_t)s * S::STAGE_BYTES;
                if (PMODE == PM_PLAIN) {
                    // fp32 rows of this chunk from the raw ring (bulk-copied by the loader warp), then free the slot
                    constexpr int RS = S::RAW_STAGES > 0 ? S::RAW_STAGES : 1;
                    const int slot = it % RS;
#ifdef BGNN_AGG_TIMING
                    long long ut_prev = clock64();
#endif
                    mbar_wait(&bars[BAR_RFULL + slot], (it / RS) & 1);
#ifdef BGNN_AGG_TIMING
                    if (ptid == 0 && grp == 0) UT_TICK(5);
#endif
                    const float* rsrc = reinterpret_cast<const float*>(raw + (size_t)slot * S::RAW_BYTES);
                    float4 av[A_F4];
#pragma unroll
                    for (int i = 0; i < A_F4; ++i) {
                        const int f = ptid + i * PROD_GROUP_THREADS;           // (row, 16 B column) of the [SUPER_M x 32] chunk
                        av[i] = *reinterpret_cast<const float4*>(rsrc + f * 4);
                    }
                    if (it >= (uint32_t)S::NSTAGE) mbar_wait(&bars[BAR_EMPTY + s], ((it / S::NSTAGE) - 1) & 1);
#ifdef BGNN_AGG_TIMING
                    if (ptid == 0 && grp == 0) UT_TICK(6);
#endif
#pragma unroll
                    for (int i = 0; i < A_F4; ++i) {                   // 2-way fp16 split -> swizzled STS
                        const int f = ptid + i * PROD_GROUP_THREADS;
                        const int mt = f >> 10, rem = f & 1023, r = rem >> 3, c4 = rem & 7;
                        uint2 hi, lo;
                        split4h(av[i], sA, hi, lo);
                        unsigned char* t0 = stage + (size_t)(mt * NSPLIT) * A_TILE_BYTES + sw64_off(r, c4 >> 1) + (c4 & 1) * 8;
                        *reinterpret_cast<uint2*>(t0) = hi;
                        *reinterpret_cast<uint2*>(t0 + A_TILE_BYTES) = lo;
                    }
                    // Release the raw slot only AFTER the stores that consumed the loaded values: an arrive issued right
                    // behind the LDS does not wait for their data, and the next TMA copy then overwrites rows that are still
                    // being read (seen as scattered wrong rows under load).
                    mbar_arrive(&bars[BAR_REMPTY + slot]);
#ifdef BGNN_AGG_TIMING
                    if (ptid == 0 && grp == 0) UT_TICK(7);
#endif
                } else {
                    // thread <-> node: its F inputs are loaded once, then the 32 hidden units of this chunk are
                    // relu(b0 + x W0^T) straight into the split tiles (W0 rows are warp-uniform broadcasts from smem)
#ifdef BGNN_AGG_TIMING
                    long long ut_prev = clock64();
#endif
                    float xr[MT][MLP0_FP];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const int64_t gr = row0 + mt * TILE_M + ptid;
#pragma unroll
                        for (int q = 0; q < MLP0_FP; ++q) xr[mt][q] = (gr < M && q < P.F) ? __ldg(P.x + gr * P.F + q) : 0.f;
                    }
                    if (it >= (uint32_t)S::NSTAGE) mbar_wait(&bars[BAR_EMPTY + s], ((it / S::NSTAGE) - 1) & 1);
#ifdef BGNN_AGG_TIMING
                    if (ptid == 0 && grp == 0) UT_TICK(6);
#endif
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
                            uint2 hi, lo;
                            split4h(make_float4(o[0], o[1], o[2], o[3]), sA, hi, lo);
                            unsigned char* t0 = stage + (size_t)(mt * NSPLIT) * A_TILE_BYTES + sw64_off(ptid, c4 >> 1) + (c4 & 1) * 8;
                            *reinterpret_cast<uint2*>(t0) = hi;
                            *reinterpret_cast<uint2*>(t0 + A_TILE_BYTES) = lo;
                        }
                    }
#ifdef BGNN_AGG_TIMING
                    if (ptid == 0 && grp == 0) UT_TICK(7);
#endif
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy writes -> async proxy (tensor core)
                mbar_arrive(&bars[BAR_FULL + s]);
            }
        }
        }
    } else if (warp == LOAD_WARP) {
        // ===================== A LOADER (one thread, PM_PLAIN): A [SUPER_M x 32] fp32 -> raw ring, ONE TMA tensor copy per chunk
        // (rows past M are zero-filled by the tensor map's bounds check) =====================
        if (PMODE == PM_SPLIT && lane == 0) {
            // A was written as fp16 hi / lo planes [M, K] by the previous kernel: two SWIZZLE_64B tensor copies per chunk land
            // in the K-major layout the MMA descriptors expect ([split][mt] tiles of 128 rows x 64 B)
            uint32_t it = 0;
            for (int64_t st = st_first; st < n_super; st += st_step) {
                const int row0 = (int)(st * SUPER_M);
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int sidx = it % S::NSTAGE;
                    if (it >= (uint32_t)S::NSTAGE) mbar_wait(&bars[BAR_EMPTY + sidx], ((it / S::NSTAGE) - 1) & 1);
                    const uint32_t bar = smem_u32(&bars[BAR_FULL + sidx]);
                    const uint32_t dst = smem_u32(base + (size_t)sidx * S::STAGE_BYTES);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)S::STAGE_BYTES) : "memory");
                    if (mc) {
                        // this CTA fetches ONE plane (rank 0: hi, rank 1: lo) and multicasts it to the same stage of both CTAs;
                        // each CTA's full barrier collects the bytes of both planes
                        const uint16_t both = 3;
                        if (nh == 0)
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;\n"
                                         ::"r"(dst), "l"(&tmapA), "r"(kc * KC), "r"(row0), "r"(bar), "h"(both) : "memory");
                        else
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;\n"
                                         ::"r"(dst + (uint32_t)(MT * A_TILE_BYTES)), "l"(&tmapA2), "r"(kc * KC), "r"(row0), "r"(bar), "h"(both) : "memory");
                        continue;
                    }
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                                 ::"r"(dst), "l"(&tmapA), "r"(kc * KC), "r"(row0), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                                 ::"r"(dst + (uint32_t)(MT * A_TILE_BYTES)), "l"(&tmapA2), "r"(kc * KC), "r"(row0), "r"(bar) : "memory");
                }
            }
        }
        if (PMODE == PM_PLAIN && lane == 0) {
            constexpr int RS = S::RAW_STAGES > 0 ? S::RAW_STAGES : 1;
            uint32_t it = 0;
            for (int64_t st = st_first; st < n_super; st += st_step) {
                const int row0 = (int)(st * SUPER_M);
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int slot = it % RS;
                    if (it >= (uint32_t)RS) mbar_wait(&bars[BAR_REMPTY + slot], ((it / RS) - 1) & 1);
                    const uint32_t bar = smem_u32(&bars[BAR_RFULL + slot]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)S::RAW_BYTES) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                                 ::"r"(smem_u32(raw + (size_t)slot * S::RAW_BYTES)), "l"(&tmapA), "r"(kc * KC), "r"(row0), "r"(bar)
                                 : "memory");
                }
            }
        }
    } else if (warp == WLOAD_WARP) {
        // ===================== W LOADER (one thread): the host pre-arranged [kc][split] tiles in the swizzled smem layout:
        // ONE bulk copy per chunk into its own ring (an own thread: its waits must not hold up the A loads) =====================
        if (lane == 0) {
            constexpr uint32_t W_BYTES = S::W_SLOT_BYTES;
            if (NH > 1) {
                // resident W: rows [nh*N, (nh+1)*N) of every K-major swizzled [NT x 32] tile are one contiguous block (8-row
                // groups); all chunks complete on ONE barrier (tx count = nkc * W_SLOT_BYTES <= 128 KB)
                const uint32_t wbar = smem_u32(&bars[BAR_WFULL]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(wbar), "r"((uint32_t)nkc * W_BYTES) : "memory");
                for (int kc = 0; kc < nkc; ++kc)
#pragma unroll
                    for (int sp = 0; sp < NSPLIT; ++sp)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                                     ::"r"(smem_u32(wring + (size_t)kc * S::W_SLOT_BYTES + (size_t)sp * S::B_TILE_BYTES)),
                                       "l"(reinterpret_cast<const unsigned char*>(P.Wsplit) +
                                           ((size_t)(kc * NSPLIT + sp) * NH + nh) * S::B_TILE_BYTES),
                                       "r"((uint32_t)S::B_TILE_BYTES), "r"(wbar)
                                     : "memory");
            } else {
                uint32_t it = 0;
                for (int64_t st = st_first; st < n_super; st += st_step) {
                    for (int kc = 0; kc < nkc; ++kc, ++it) {
                        const int w = it % NW;
                        if (it >= (uint32_t)NW) mbar_wait(&bars[BAR_WEMPTY + w], ((it / NW) - 1) & 1);
                        const uint32_t wbar = smem_u32(&bars[BAR_WFULL + w]);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(wbar), "r"(W_BYTES) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                                     ::"r"(smem_u32(wring + (size_t)w * S::W_SLOT_BYTES)),
                                       "l"(reinterpret_cast<const unsigned char*>(P.Wsplit) + (size_t)kc * W_BYTES), "r"(W_BYTES), "r"(wbar)
                                     : "memory");
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ===================== MMA ISSUER (one thread) =====================
        if (lane == 0) {
            // instruction descriptor: D = F32 (bit 4), A = B = F16 (format 0), both K-major, N, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            uint32_t it = 0, t = 0;
            UT_DECL();
#ifdef BGNN_AGG_TIMING
            const long long ut_c0 = clock64();
            unsigned long long ut_g0;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ut_g0));
#endif
            if (NH > 1) mbar_wait(&bars[BAR_WFULL], 0);                // resident W has landed
            for (int64_t st = st_first; st < n_super; st += st_step, ++t) {
                const uint32_t as = t % S::ACC_STAGES;
                if (t >= (uint32_t)S::ACC_STAGES) mbar_wait(&bars[BAR_TEMPTY + as], (t / S::ACC_STAGES - 1) & 1);   // epilogue drained this stage
                UT_TICK(0);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int s = it % S::NSTAGE, w = NH > 1 ? kc : (int)(it % NW);
                    if (NH == 1) mbar_wait(&bars[BAR_WFULL + w], (it / NW) & 1);
                    mbar_wait(&bars[BAR_FULL + s], (it / S::NSTAGE) & 1);
                    UT_TICK(1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t sA_ = smem_u32(base + (size_t)s * S::STAGE_BYTES), sB_ = smem_u32(wring + (size_t)w * S::W_SLOT_BYTES);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t d_main = tmem_base + as * S::ACC_COLS + (uint32_t)(mt * 2 * N), d_corr = d_main + (uint32_t)N;
                        // p = 0: a_hi*b_hi -> main; p = 1, 2: a_lo*b_hi, a_hi*b_lo (both carry the 2^11 factor) -> corr
                        const int pa[3] = {0, 1, 0}, pb[3] = {0, 0, 1};
#pragma unroll
                        for (int p = 0; p < 3; ++p) {
#pragma unroll
                            for (int ks = 0; ks < KC / 16; ++ks) {
                                const uint64_t da = make_desc(sA_ + (PMODE == PM_SPLIT ? pa[p] * MT + mt : mt * NSPLIT + pa[p]) * A_TILE_BYTES + ks * 32);
                                const uint64_t db = make_desc(sB_ + pb[p] * S::B_TILE_BYTES + ks * 32);
                                const bool first = kc == 0 && ks == 0 && (p == 0 || p == 1);
                                umma_f16(p == 0 ? d_main : d_corr, da, db, idesc, first ? 0u : 1u);
                            }
                        }
                    }
                    if (mc) umma_commit_mc(&bars[BAR_EMPTY + s], 3);   // ... in BOTH CTAs of the pair (either may refill the stage)
                    else umma_commit(&bars[BAR_EMPTY + s]);            // A stage and W slot free when these MMAs retire
                    if (NH == 1) umma_commit(&bars[BAR_WEMPTY + w]);
                    UT_TICK(2);
                }
                umma_commit(&bars[BAR_TFULL + as]);                    // accumulators complete
            }
#ifdef BGNN_AGG_TIMING
            {
                unsigned long long ut_g1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ut_g1));
                if (BGNN_UT_FILTER) {
                    atomicAdd(&g_umma_t[8], (unsigned long long)(clock64() - ut_c0));
                    atomicAdd(&g_umma_t[9], ut_g1 - ut_g0);
                }
            }
#endif
        }
    } else if ((warp >> 2) < S::EPI_SPLIT) {
        // ===================== EPILOGUE: warp w <-> TMEM lanes 32*(w%4) .., one thread per output row =====================
        const int quarter = warp & 3, half = warp >> 2;
        constexpr int NCOL = N / S::EPI_SPLIT;                 // columns owned by this warp
        const int cbeg = half * NCOL;
        float cmax = 0.f;
        uint32_t t = 0;
        for (int64_t st = st_first; st < n_super; st += st_step, ++t) {
            const int64_t row0 = st * SUPER_M;
            const uint32_t as = t % S::ACC_STAGES;
#ifdef BGNN_AGG_TIMING
            long long ut_prev = clock64();
#endif
            mbar_wait(&bars[BAR_TFULL + as], (t / S::ACC_STAGES) & 1);
#ifdef BGNN_AGG_TIMING
            if (warp == 0 && lane == 0) UT_TICK(3);
#endif
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int64_t row = row0 + mt * TILE_M + quarter * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * S::ACC_COLS + (uint32_t)(mt * 2 * N);
                if (EMODE == EM_PROJ) {
                    float as_acc = 0.f, ad_acc = 0.f;
#pragma unroll 1
                    for (int cl = cbeg; cl < cbeg + NCOL; cl += 32) {
                        const int c0 = cg0 + cl;               // global output column (cl: column inside the accumulator)
                        const int pc = NH > 1 ? cl : c0;       // column inside this CTA's epilogue parameters
                        float x[32];
                        tmem_ld_sum32(taddr + (uint32_t)cl, taddr + (uint32_t)(N + cl), out_scale, x);
                        if (P.bias) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
#if BGNN_EPI_PACKED
                                const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(s_par + 2 * S::PARW + pc + j);
                                unpk2(fadd2(pk2(x[j], x[j + 1]), b.x), x[j], x[j + 1]);
                                unpk2(fadd2(pk2(x[j + 2], x[j + 3]), b.y), x[j + 2], x[j + 3]);
#else
                                const float4 b = *reinterpret_cast<const float4*>(s_par + 2 * S::PARW + pc + j);
                                x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
#endif
                            }
                        }
                        if (P.cmax_out) {
                            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                m0 = fmaxf(m0, fabsf(x[j])); m1 = fmaxf(m1, fabsf(x[j + 1]));
                                m2 = fmaxf(m2, fabsf(x[j + 2])); m3 = fmaxf(m3, fabsf(x[j + 3]));
                            }
                            cmax = fmaxf(cmax, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                        }
                        if (P.alpha) {
                            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#if BGNN_EPI_PACKED
                            uint64_t s01 = 0ull, s23 = 0ull, d01 = 0ull, d23 = 0ull;       // the same four partial sums, as pairs
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const ulonglong2 s4 = *reinterpret_cast<const ulonglong2*>(s_par + pc + j);
                                const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(s_par + S::PARW + pc + j);
                                const uint64_t x01 = pk2(x[j], x[j + 1]), x23 = pk2(x[j + 2], x[j + 3]);
                                s01 = ffma2(x01, s4.x, s01); s23 = ffma2(x23, s4.y, s23);
                                d01 = ffma2(x01, d4.x, d01); d23 = ffma2(x23, d4.y, d23);
                            }
                            unpk2(s01, s0, s1); unpk2(s23, s2, s3); unpk2(d01, d0, d1); unpk2(d23, d2, d3);
#else
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 s4 = *reinterpret_cast<const float4*>(s_par + pc + j);
                                const float4 d4 = *reinterpret_cast<const float4*>(s_par + S::PARW + pc + j);
                                s0 = fmaf(x[j], s4.x, s0); s1 = fmaf(x[j + 1], s4.y, s1);
                                s2 = fmaf(x[j + 2], s4.z, s2); s3 = fmaf(x[j + 3], s4.w, s3);
                                d0 = fmaf(x[j], d4.x, d0); d1 = fmaf(x[j + 1], d4.y, d1);
                                d2 = fmaf(x[j + 2], d4.z, d2); d3 = fmaf(x[j + 3], d4.w, d3);
                            }
#endif
                            as_acc += (s0 + s1) + (s2 + s3);
                            ad_acc += (d0 + d1) + (d2 + d3);
                            if (((c0 + 32) % P.CH) == 0) {          // a head's channels are complete
                                if (row < M) {
                                    const int h = (c0 + 32) / P.CH - 1;
                                    P.alpha[row * (2 * P.Htot) + h] = as_acc;
                                    P.alpha[row * (2 * P.Htot) + P.Htot + h] = ad_acc;
                                }
                                as_acc = 0.f; ad_acc = 0.f;
                            }
                        }
                        if (!S::TMA_STORE) {         // thread-per-row 256-bit stores (32 different lines per instruction)
                            if (row < M) {
                                float* dst = P.head_major ? P.Cout + ((int64_t)(c0 / P.CH) * M + row) * P.CH + (c0 % P.CH)
                                                          : P.Cout + row * NT + c0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) st_global_v8(dst + 8 * j, x + 8 * j);
                            }
                        } else {   // chunk [32 rows x 32 cols] -> SWIZZLE_128B staging tile -> one TMA tensor store (rows >= M are clipped)
                            unsigned char* tile = stg + (size_t)warp * 4096;
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // previous store has read the tile
                            __syncwarp();
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                *reinterpret_cast<float4*>(tile + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                                    make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {
                                const int cx = P.head_major ? (c0 % P.CH) : c0, cz = P.head_major ? (c0 / P.CH) : 0;
                                const int cy = (int)(row0 + mt * TILE_M + quarter * 32);
                                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];\n"
                                             ::"l"(&tmapC), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(tile)) : "memory");
                                asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                            }
                        }
                    }
                } else {
                    // heads: columns [b*h2, (b+1)*h2) = hidden units of head b (0 cls, 1 conf, 2 corr), h2 % 32 == 0
                    const float* w2 = s_par + N;
                    const float* b2 = w2 + HEADS_MAXOUT * HEADS_MAXH2;
                    const int h2 = P.h2, ncls = P.ncls, real = P.nblk * h2;
                    // Work split between the two epilogue halves (warps q and q + 4 share TMEM lanes 32q ..): half 0 takes the
                    // classification block (ncls dot products, softmax, argmax), half 1 the confidence and correction blocks
                    // and the action, for which it receives the class through shared memory (one 64-thread named barrier
                    // per tile; the hand-over slot alternates with the tile parity, so a slot is rewritten only after the
                    // barrier of the following tile, i.e. after its reader is done with it).
                    float lg[8], pc = 0.f, pr = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) lg[k] = 0.f;
                    int* s_arg = reinterpret_cast<int*>(s_par + N + HEADS_MAXOUT * HEADS_MAXH2 + 16) + (t & 1u) * TILE_M;
#pragma unroll 1
                    for (int c0 = 0; c0 < real; c0 += 32) {
                        const int blk = c0 / h2, u0 = c0 - blk * h2;
                        if ((blk >= 1) != (half == 1)) continue;
                        float x[32];
                        tmem_ld_sum32(taddr + (uint32_t)c0, taddr + (uint32_t)(N + c0), out_scale, x);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(s_par + c0 + j);
                            x[j] = fmaxf(x[j] + b.x, 0.f); x[j + 1] = fmaxf(x[j + 1] + b.y, 0.f);
                            x[j + 2] = fmaxf(x[j + 2] + b.z, 0.f); x[j + 3] = fmaxf(x[j + 3] + b.w, 0.f);
                        }
                        // four independent partial sums per output (a 32-long serial FMA chain is latency-bound), kept as two
                        // packed pairs: (p0, p1), (p2, p3)
                        auto dot32 = [&](const float* wv) {
                            uint64_t p01 = 0ull, p23 = 0ull;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(wv + j);
                                p01 = ffma2(pk2(x[j], x[j + 1]), w.x, p01);
                                p23 = ffma2(pk2(x[j + 2], x[j + 3]), w.y, p23);
                            }
                            float p0, p1, p2, p3;
                            unpk2(p01, p0, p1); unpk2(p23, p2, p3);
                            return (p0 + p1) + (p2 + p3);
                        };
                        if (blk == 0) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (k < ncls) lg[k] += dot32(w2 + k * h2 + u0);
                        } else {
                            const float acc = dot32(w2 + (blk == 1 ? ncls : ncls + 1) * h2 + u0);
                            if (blk == 1) pc += acc; else pr += acc;
                        }
                    }
                    if (half == 0) {
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
                                if (row < M && P.logits) P.logits[row * ncls + k] = lg[k];
                                if (row < M && P.probs) P.probs[row * ncls + k] = pb[k];
                            }
                        if (row < M && P.cls) P.cls[row] = arg;
                        s_arg[quarter * 32 + lane] = arg;
                        asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quarter) : "memory");
                    } else {
                        const float cf = 1.f / (1.f + expf(-(pc + b2[ncls])));
                        if (row < M && P.conf) P.conf[row] = cf;
                        if (row < M && P.corr && P.nblk > 2) P.corr[row] = pr + b2[ncls + 1];
                        asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quarter) : "memory");
                        if (row < M && P.action) {
                            const int arg = s_arg[quarter * 32 + lane];
                            int64_t a = (arg == 2 && cf > P.thr_auto) ? 1 : 0;          // gnn.py:439-441
                            if (cf < P.thr_review) a = 2;                                // gnn.py:444-445
                            P.action[row] = a;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            mbar_arrive(&bars[BAR_TEMPTY + as]);
#ifdef BGNN_AGG_TIMING
            if (warp == 0 && lane == 0) UT_TICK(4);
#endif
        }
        if (S::TMA_STORE && lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");     // all tensor stores done
        if (EMODE == EM_PROJ && P.cmax_out) {       // max |C| written by this launch (feeds the next kernel's operand scale)
#pragma unroll
            for (int o = 16; o; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
            if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(P.cmax_out), __float_as_uint(cmax));
        }
    }
    // teardown
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(S::TMEM_COLS) : "memory");
    }
    if (mc) cluster_sync_all();          // no CTA leaves while its peer may still arrive on its barriers
}

// =====================================================================================================================
// cta_group::2 projection (EM_PROJ epilogue, N = 256; operands PM_SPLIT with K <= 256, or PM_MLP0 with K = 64 -- see Cfg below):
// the two CTAs of a cluster own ONE 256 x 256 output tile.  CTA r stages its own 128 rows of A (fp16 plane chunks of 16 KB) and keeps its N half of the split
// weights resident (128 KB); the leader's MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16), for which each
// SM reads only its own A rows and its own B half from shared memory: half the operand wavefronts per output of the
// N = 2 x 128 multicast pair, whose tensor-core operand reads (46 % of the L1 data pipe) plus the epilogue's thread-per-row
// stores (59 %) saturated that pipe (profiles/r2c_ncu_full_summary.csv).
//   ONE accumulator per tile (256 TMEM columns, double-buffered: the epilogue of tile t overlaps the MMAs of tile t + 1).  The
//   K extent is walked twice: pass 1 accumulates the two correction products a_lo*w_hi + a_hi*w_lo (both carry the 2^11 factor
//   of the lo planes), pass 2 the main product a_hi*w_hi, whose FIRST instruction scales the accumulator by 2^-11
//   (scale-input-d = 11: D = A*B + D * 2^-11) -- the same fp32 sum as main + corr * 2^-11 without a second accumulator.  The
//   hi plane is therefore fetched twice (the second time from L2); the shared-memory fill per output still drops by 25 %.
//   A ring: 6 units of [128 rows x 64] fp16 (128-byte rows, SWIZZLE_128B; the resident W keeps its 32-wide SWIZZLE_64B chunks --
//   the two operand descriptors carry their own layout), filled by two loader threads that take alternate units.
//   epilogue (PM_SPLIT: 8 warps x 128 columns): tcgen05.ld of the next chunk is in flight while a chunk is finished; thread-per-row
//   256-bit stores (staging tiles for TMA stores would cost 32 KB = 2 ring units: 2.24 vs 1.65 ms; PM_MLP0 has the room).  The optional
//   bias is read from global memory (warp-uniform, L1-resident): its 1 KB of shared memory is what lets the sixth unit fit.
//   measured (2048^2 tile, M = 4.1 M, K = N = 256): 1.53 ms vs 1.78 ms for the N = 2 x 128 multicast pair on the same box (5 ring
//   units: 1.65 ms; L2 prefetches of the pair's next tile: 1.56 ms -- the ring depth matters, the HBM -> L2 leg does not).
//   barriers (same offsets in both CTAs; "leader" = cluster rank 0):
//     FULL[u]    leader's.  PM_SPLIT: 1 arrival -- the leader's loader arms it for the bytes of BOTH CTAs (arrive.expect_tx 32 KB), the
//                TMA copies of both CTAs complete on it (cp.async.bulk.tensor ... cta_group::2).  PM_MLP0: 2 arrivals -- one
//                elected producer thread per CTA after the producers' own named barrier (remote arrive for the peer)
//     EMPTY[u]   own, 1 arrival: the leader's tcgen05.commit multicast to both CTAs when the MMAs reading the unit have retired
//     TFULL[a]   own, 1 arrival: commit multicast after the last MMA of a tile
//     TEMPTY[a]  leader's, 2 x EPW arrivals: lane 0 of every epilogue warp of both CTAs after the warp's last tcgen05.ld of the tile
//     WFULL      own (resident W landed); WPEER leader's: the peer's W loader arrives once its own WFULL completed
namespace pairk {
constexpr int PN = 256, PHALF = 128;
constexpr int KCA = 64;                            // K elements per A unit: 128-byte rows (SWIZZLE_128B), half the TMA row requests
                                                   // of 64-byte rows
constexpr int UNIT_B = TILE_M * KCA * 2;           // 16 KB: ONE plane chunk [128 rows x 64] fp16
constexpr int BT_B = PHALF * KC * 2;               // 8 KB: one split of this CTA's N half of a 32-wide W chunk
constexpr int WSLOT_B = NSPLIT * BT_B;
// MODE = PM_SPLIT: operands arrive as fp16 planes (two TMA loader threads); K <= 256, the hi plane is fetched for both passes
// MODE = PM_MLP0:  eight producer warps compute this CTA's 128 rows of relu(x W0^T + b0) straight into a {hi, lo} unit pair
//                  (K = 64: one unit per plane holds the whole K extent and serves both passes); three tile slots; the
//                  spare shared memory goes to staging tiles for TMA tensor stores
template <int MODE> struct Cfg {
    static constexpr bool MLP = MODE == PM_MLP0;
#ifndef BGNN_PAIR_MLP_PRODW
#define BGNN_PAIR_MLP_PRODW 8          /* producer warps of the MLP mode: 4 (thread <-> row) or 8 (thread <-> row, half of the hidden units: 1.01 vs 1.03 ms) */
#endif
    static constexpr int FEED_WARPS = MLP ? BGNN_PAIR_MLP_PRODW : 2;
#ifndef BGNN_PAIR_EPW_MLP
#define BGNN_PAIR_EPW_MLP 16
#endif
    // epilogue warps: warp w <-> TMEM lanes 32 (w % 4).., columns [(w / 4) * 1024 / EPW, +1024 / EPW).  MLP: the epilogue IS the kernel
    // (K = 64: 12 MMAs per tile), a chain of dependent instructions per warp -> four warps per scheduler instead of two
#ifndef BGNN_PAIR_EPW_SPLIT
#define BGNN_PAIR_EPW_SPLIT 8
#endif
    static constexpr int EPW = MLP ? BGNN_PAIR_EPW_MLP : BGNN_PAIR_EPW_SPLIT;
    static constexpr int NCOLW = 4 * PN / EPW;                              // columns per epilogue warp: 128 or 64
    static constexpr int THREADS = (EPW + 2 + FEED_WARPS) * 32;            // epilogue | MMA issuer | W loader | feeders
    static constexpr int NU = 6;                                            // A ring (MLP: 3 tile slots x {hi, lo})
    static constexpr int MAXKC = MLP ? 2 : 8;                               // resident W chunks of 32
    static constexpr int STG_B = MLP ? EPW * 4096 : 0;                      // [32 x 32] fp32 staging tile per epilogue warp
    static constexpr int PROD_F = MLP ? KCA * MLP0_FP + KCA : 0;            // W0 [64][8] | b0 [64]
    static constexpr size_t BYTES = (size_t)NU * UNIT_B + (size_t)MAXKC * WSLOT_B + STG_B + (2 * PN + PROD_F) * sizeof(float) + 32 * 8 + 16;
    static_assert(BYTES <= 232448, "shared memory budget exceeded");
};
enum { B_FULL = 0, B_EMPTY = 6, B_TFULL = 12, B_TEMPTY = 14, B_WFULL = 16, B_WPEER, B_NUM };
}  // namespace pairk

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// (relaxed: a release at cluster scope costs a full memory fence per call -- measured 1 300 cycles per fetch in the loader
//  thread -- and nothing is handed over through generic memory here: the data arrives with the TMA's own complete_tx)
__device__ __forceinline__ void mbar_expect_tx_remote(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;\n" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// default semantics (release at CTA scope): orders this CTA's own shared-memory writes, which its own SM's tensor core reads
__device__ __forceinline__ void mbar_arrive_remote_cta(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {       // arrivals may come from the peer CTA
    uint32_t ok;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (128-byte rows, 1024 B between 8-row groups; layout type 2)
__device__ __forceinline__ uint64_t make_desc128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D = A*B + D * 2^-11: folds the correction accumulator (which carries the 2^11 factor of the lo planes) into the main sum
__device__ __forceinline__ void umma_f16_cg2_fold(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p, 11;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t addr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
}

#ifndef BGNN_EPI_SKIP
#define BGNN_EPI_SKIP 0          /* dev: bit 0 skip alpha, bit 1 skip stores, bit 2 skip bias + cmax (timing experiments) */
#endif
template <int MODE>
__global__ void __launch_bounds__(pairk::Cfg<MODE>::THREADS, 1)
proj_pair_kernel(const __grid_constant__ UmmaArgs P, const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapA2,
                 const __grid_constant__ CUtensorMap tmapC) {
    using namespace pairk;
    using G = Cfg<MODE>;
    constexpr bool MLP = G::MLP;
    constexpr int NU = G::NU, THREADS = G::THREADS;
    constexpr int EPW = G::EPW, MMA_W = EPW, W_WARP = EPW + 1, FEED_WARP0 = EPW + 2;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* units = smem_raw;                                               // [NU] plane chunks of this CTA's 128 rows
    unsigned char* wres = units + (size_t)NU * UNIT_B;                             // [nkc][hi half tile | lo half tile]
    unsigned char* stg = wres + (size_t)G::MAXKC * WSLOT_B;                        // MLP: [8 warps][32 x 32 fp32, SWIZZLE_128B]
    float* s_par = reinterpret_cast<float*>(stg + G::STG_B);                       // att_src | att_dst   [2][256]
    float* s_prod = s_par + 2 * PN;                                                // MLP: W0 [64][8] (F padded with zeros) | b0 [64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_prod + G::PROD_F);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + B_NUM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
    const int64_t M = P.M;
    const int nkc = P.K / KC, nkca = P.K / KCA;
    const int64_t n_tiles = (M + 2 * TILE_M - 1) / (2 * TILE_M);
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    // power-of-two operand scale from the upper bound of max|A| (PM_SPLIT: applied by the kernel that wrote the planes)
    float sA = 1.f, out_scale = P.w_inv_scale;
    {
        float bound = P.amax_add;
        if (P.amax_src) bound = fmaf(P.amax_mul, __ldg(P.amax_src), P.amax_add);
        const int e = (int)((__float_as_uint(bound) >> 23) & 0xffu);
        const int es = 253 - e;
        if (e > 0 && e < 255 && es >= 1 && es <= 253) {
            sA = __uint_as_float((uint32_t)es << 23);
            out_scale *= __uint_as_float((uint32_t)(254 - es) << 23);
        }
    }
    if (tid == 0) {
        // MLP: FULL[slot] collects one arrival per CTA (an elected producer thread, after the producers' own barrier)
        for (int u = 0; u < NU; ++u) { mbar_init(&bars[B_FULL + u], MLP ? 2 : 1); mbar_init(&bars[B_EMPTY + u], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bars[B_TFULL + a], 1); mbar_init(&bars[B_TEMPTY + a], 2 * EPW); }
        mbar_init(&bars[B_WFULL], 1);
        mbar_init(&bars[B_WPEER], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < PN; i += THREADS) {
        s_par[i] = P.alpha ? __ldg(P.att_src + i) : 0.f;
        s_par[PN + i] = P.alpha ? __ldg(P.att_dst + i) : 0.f;
    }
    if (MLP) {
        for (int i = tid; i < KCA * MLP0_FP; i += THREADS) {
            const int kk = i / MLP0_FP, f = i % MLP0_FP;
            s_prod[i] = f < P.F ? __ldg(P.W0 + kk * P.F + f) : 0.f;
        }
        for (int i = tid; i < KCA; i += THREADS) s_prod[KCA * MLP0_FP + i] = __ldg(P.b0 + i);
    }
    if (warp == MMA_W) {                       // the same warp of both CTAs allocates all 512 columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    cluster_sync_all();                            // the peer's barriers are initialised before anything arrives on them

    if (warp >= FEED_WARP0) {
        if constexpr (MLP) {
        // ===================== PRODUCERS (128 threads per CTA, thread <-> own row): relu(b0 + x W0^T) * sA -> fp16 hi / lo, written
        // as 16-byte chunks of the K-major SWIZZLE_128B units (chunk c of row r at (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4))
        const int ptid = tid - FEED_WARP0 * 32;
        const int r = ptid & (TILE_M - 1), chalf = ptid >> 7;                    // (8 producer warps: two threads share a row)
        constexpr int NCH = (KCA / 8) / (G::FEED_WARPS / 4);                     // 16-byte chunks (8 hidden units) per thread
        const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        const uint32_t full0 = mapa_u32(smem_u32(&bars[B_FULL]), 0);
        uint32_t t = 0;
        for (int64_t tile = pair; tile < n_tiles; tile += npairs, ++t) {
            const int slot = t % 3;
            const int64_t gr = tile * 2 * TILE_M + rank * TILE_M + r;
            float xr[MLP0_FP];
#pragma unroll
            for (int q = 0; q < MLP0_FP; ++q) xr[q] = (gr < M && q < P.F) ? __ldg(P.x + gr * P.F + q) : 0.f;
#ifdef BGNN_AGG_TIMING
            long long ut_prev = clock64();
#endif
            if (t >= 3) mbar_wait(&bars[B_EMPTY + slot], ((t / 3) - 1) & 1);
#ifdef BGNN_AGG_TIMING
            if (ptid == 0 && rank == 0) UT_TICKP(5);
#endif
            unsigned char* u_hi = units + (size_t)(2 * slot) * UNIT_B + rowoff;
            unsigned char* u_lo = u_hi + UNIT_B;
#pragma unroll 2
            for (int c = chalf * NCH; c < (chalf + 1) * NCH; ++c) {
                float o[8];
#pragma unroll
                for (int uu = 0; uu < 8; ++uu) {
                    const float4 w0 = *reinterpret_cast<const float4*>(s_prod + (c * 8 + uu) * MLP0_FP);
                    const float4 w1 = *reinterpret_cast<const float4*>(s_prod + (c * 8 + uu) * MLP0_FP + 4);
                    float acc = xr[0] * w0.x;
                    acc = fmaf(xr[1], w0.y, acc); acc = fmaf(xr[2], w0.z, acc); acc = fmaf(xr[3], w0.w, acc);
                    acc = fmaf(xr[4], w1.x, acc); acc = fmaf(xr[5], w1.y, acc); acc = fmaf(xr[6], w1.z, acc);
                    acc = fmaf(xr[7], w1.w, acc);
                    o[uu] = fmaxf(acc + s_prod[KCA * MLP0_FP + c * 8 + uu], 0.f);
                }
                uint2 h0, l0, h1, l1;
                split4h(make_float4(o[0], o[1], o[2], o[3]), sA, h0, l0);
                split4h(make_float4(o[4], o[5], o[6], o[7]), sA, h1, l1);
                const uint32_t phys = (uint32_t)((c ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(u_hi + phys) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                *reinterpret_cast<uint4*>(u_lo + phys) = make_uint4(l0.x, l0.y, l1.x, l1.y);
            }
#ifdef BGNN_AGG_TIMING
            if (ptid == 0 && rank == 0) UT_TICKP(6);
#endif
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy writes -> async proxy (tensor core)
            asm volatile("bar.sync 1, %0;\n" ::"n"(G::FEED_WARPS * 32) : "memory");      // all rows of this CTA's units are written
            if (ptid == 0) mbar_arrive_remote_cta(full0 + 8u * slot);                // one arrival per tile and CTA
#ifdef BGNN_AGG_TIMING
            if (ptid == 0 && rank == 0) UT_TICKP(7);
#endif
        }
        } else {
        // ===================== A LOADERS (two threads per CTA, alternate units): own 128 rows; pass 1: hi, lo of every chunk;
        // pass 2: hi again.  The LEADER's thread arms the unit's barrier for the bytes of both CTAs (the peer's copy may complete
        // first: the transaction count then goes negative until the leader's arrive.expect_tx, the phase cannot complete before it)
        if (lane == 0 && warp < FEED_WARP0 + 2) {
            const uint32_t me = (uint32_t)(warp - FEED_WARP0);
            const uint32_t full0 = mapa_u32(smem_u32(&bars[B_FULL]), 0);
            uint32_t u = 0;
            UT_DECL();
            auto fetch = [&](const CUtensorMap* tm, int kc, int row0) {
                if ((u & 1u) == me) {
                    const int sl = u % NU;
                    if (rank == 0 && me == 0) UT_TICKP(6);
                    if (u >= (uint32_t)NU) mbar_wait(&bars[B_EMPTY + sl], ((u / NU) - 1) & 1);
                    if (rank == 0 && me == 0) UT_TICKP(5);
                    if (rank == 0)
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bars[B_FULL + sl])), "r"(2u * UNIT_B) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                                 ::"r"(smem_u32(units + (size_t)sl * UNIT_B)), "l"(tm), "r"(kc * KCA), "r"(row0), "r"(full0 + 8u * sl) : "memory");
                }
                ++u;
            };
            for (int64_t tile = pair; tile < n_tiles; tile += npairs) {
                const int row0 = (int)(tile * 2 * TILE_M + rank * TILE_M);
                for (int kc = 0; kc < nkca; ++kc) { fetch(&tmapA, kc, row0); fetch(&tmapA2, kc, row0); }
                for (int kc = 0; kc < nkca; ++kc) fetch(&tmapA, kc, row0);
            }
        }
        }
    } else if (warp == W_WARP) {
        // ===================== W LOADER (one thread per CTA): this CTA's N half of every chunk, once =====================
        if (lane == 0) {
            const uint32_t wbar = smem_u32(&bars[B_WFULL]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(wbar), "r"((uint32_t)(nkc * WSLOT_B)) : "memory");
            for (int kc = 0; kc < nkc; ++kc)
#pragma unroll
                for (int sp = 0; sp < NSPLIT; ++sp)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                                 ::"r"(smem_u32(wres + (size_t)kc * WSLOT_B + (size_t)sp * BT_B)),
                                   "l"(reinterpret_cast<const unsigned char*>(P.Wsplit) + ((size_t)(kc * NSPLIT + sp) * 2 + rank) * BT_B),
                                   "r"((uint32_t)BT_B), "r"(wbar)
                                 : "memory");
            if (rank != 0) {
                mbar_wait(&bars[B_WFULL], 0);
                mbar_arrive_remote(mapa_u32(smem_u32(&bars[B_WPEER]), 0));
            }
        }
    } else if (warp == MMA_W) {
        // ===================== MMA ISSUER (one thread of the leader CTA) =====================
        if (lane == 0 && rank == 0) {
            // instruction descriptor: D = F32, A = B = F16, both K-major, N = 256, M = 256 (128 rows per CTA)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(PN >> 3) << 17) | ((uint32_t)((2 * TILE_M) >> 4) << 24);
            mbar_wait(&bars[B_WFULL], 0);
            mbar_wait_cluster(&bars[B_WPEER], 0);
            uint32_t u = 0, t = 0;
            UT_DECL();
#ifdef BGNN_AGG_TIMING
            const long long ut_c0 = clock64();
            unsigned long long ut_g0;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ut_g0));
#endif
            for (int64_t tile = pair; tile < n_tiles; tile += npairs, ++t) {
                const uint32_t as = t & 1u;
                if (t >= 2) mbar_wait_cluster(&bars[B_TEMPTY + as], ((t >> 1) - 1) & 1);          // both epilogues drained this accumulator
                UT_TICKP(0);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t d = tmem_base + as * (uint32_t)PN;
                if constexpr (MLP) {
                    // K = 64: the tile's {hi, lo} unit pair holds the whole K extent and serves both passes
                    const int slot = t % 3;
                    mbar_wait_cluster(&bars[B_FULL + slot], (t / 3) & 1);
                    UT_TICKP(1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t a_hi = smem_u32(units + (size_t)(2 * slot) * UNIT_B), a_lo = a_hi + (uint32_t)UNIT_B;
                    const uint32_t b0 = smem_u32(wres);
#pragma unroll
                    for (int ks = 0; ks < KCA / 16; ++ks)
                        umma_f16_cg2(d, make_desc128(a_lo + ks * 32), make_desc(b0 + (uint32_t)((ks >> 1) * WSLOT_B + (ks & 1) * 32)), idesc, ks == 0 ? 0u : 1u);
#pragma unroll
                    for (int ks = 0; ks < KCA / 16; ++ks)
                        umma_f16_cg2(d, make_desc128(a_hi + ks * 32), make_desc(b0 + (uint32_t)((ks >> 1) * WSLOT_B + BT_B + (ks & 1) * 32)), idesc, 1u);
#pragma unroll
                    for (int ks = 0; ks < KCA / 16; ++ks) {
                        const uint32_t b_hi = b0 + (uint32_t)((ks >> 1) * WSLOT_B + (ks & 1) * 32);
                        if (ks == 0) umma_f16_cg2_fold(d, make_desc128(a_hi), make_desc(b_hi), idesc);
                        else umma_f16_cg2(d, make_desc128(a_hi + ks * 32), make_desc(b_hi), idesc, 1u);
                    }
                    umma_commit_cg2(&bars[B_EMPTY + slot], 3);       // the slot is free in BOTH CTAs when these MMAs retire
                    UT_TICKP(2);
                } else {
                // pass 1: corr = a_lo * w_hi + a_hi * w_lo   (an A unit spans two 32-wide W chunks: k16 step ks <-> chunk 2 kca + ks / 2)
                for (int kca = 0; kca < nkca; ++kca, u += 2) {
                    const int s_hi = u % NU, s_lo = (u + 1) % NU;
                    mbar_wait_cluster(&bars[B_FULL + s_hi], (u / NU) & 1);
                    mbar_wait_cluster(&bars[B_FULL + s_lo], ((u + 1) / NU) & 1);
                    UT_TICKP(1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t a_hi = smem_u32(units + (size_t)s_hi * UNIT_B), a_lo = smem_u32(units + (size_t)s_lo * UNIT_B);
                    const uint32_t b0 = smem_u32(wres + (size_t)(2 * kca) * WSLOT_B);
#pragma unroll
                    for (int ks = 0; ks < KCA / 16; ++ks) {
                        const uint32_t b_hi = b0 + (uint32_t)((ks >> 1) * WSLOT_B + (ks & 1) * 32);
                        umma_f16_cg2(d, make_desc128(a_lo + ks * 32), make_desc(b_hi), idesc, (kca == 0 && ks == 0) ? 0u : 1u);
                    }
#pragma unroll
                    for (int ks = 0; ks < KCA / 16; ++ks) {
                        const uint32_t b_lo = b0 + (uint32_t)((ks >> 1) * WSLOT_B + BT_B + (ks & 1) * 32);
                        umma_f16_cg2(d, make_desc128(a_hi + ks * 32), make_desc(b_lo), idesc, 1u);
                    }
                    umma_commit_cg2(&bars[B_EMPTY + s_hi], 3);       // the units are free in BOTH CTAs when these MMAs retire
                    umma_commit_cg2(&bars[B_EMPTY + s_lo], 3);
                    UT_TICKP(2);
                }
                // pass 2: main = a_hi * w_hi on top of corr * 2^-11
                for (int kca = 0; kca < nkca; ++kca, ++u) {
                    const int s_hi = u % NU;
                    mbar_wait_cluster(&bars[B_FULL + s_hi], (u / NU) & 1);
                    UT_TICKP(7);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t a_hi = smem_u32(units + (size_t)s_hi * UNIT_B);
                    const uint32_t b0 = smem_u32(wres + (size_t)(2 * kca) * WSLOT_B);
#pragma unroll
                    for (int ks = 0; ks < KCA / 16; ++ks) {
                        const uint32_t b_hi = b0 + (uint32_t)((ks >> 1) * WSLOT_B + (ks & 1) * 32);
                        if (kca == 0 && ks == 0) umma_f16_cg2_fold(d, make_desc128(a_hi), make_desc(b_hi), idesc);
                        else umma_f16_cg2(d, make_desc128(a_hi + ks * 32), make_desc(b_hi), idesc, 1u);
                    }
                    umma_commit_cg2(&bars[B_EMPTY + s_hi], 3);
                    UT_TICKP(2);
                }
                }
                umma_commit_cg2(&bars[B_TFULL + as], 3);             // accumulator complete (both CTAs' epilogues)
            }
#ifdef BGNN_AGG_TIMING
            {
                unsigned long long ut_g1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ut_g1));
                atomicAdd(&g_umma_t[8], (unsigned long long)(clock64() - ut_c0));
                atomicAdd(&g_umma_t[9], ut_g1 - ut_g0);
            }
#endif
        }
    } else {
        // ===================== EPILOGUE (both CTAs): warp w <-> TMEM lanes 32*(w%4).., columns [128*(w/4), +128) =====================
        const int quarter = warp & 3;
        const int cbeg = (warp >> 2) * G::NCOLW;
        float cmax = 0.f;
        uint32_t t = 0;
        for (int64_t tile = pair; tile < n_tiles; tile += npairs, ++t) {
            const uint32_t as = t & 1u;
#ifdef BGNN_AGG_TIMING
            long long ut_prev = clock64();
#endif
            mbar_wait(&bars[B_TFULL + as], (t >> 1) & 1);
#ifdef BGNN_AGG_TIMING
            if (warp == 0 && lane == 0 && rank == 0) UT_TICKP(3);
#endif
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int64_t row = tile * 2 * TILE_M + rank * TILE_M + quarter * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * (uint32_t)PN + (uint32_t)cbeg;
            float as_acc = 0.f, ad_acc = 0.f;
            uint32_t v[32];
            constexpr bool PREFETCH = G::EPW == 8;        // two warps per scheduler: keep the next chunk's tcgen05.ld in flight
            if (PREFETCH) tmem_ld32_issue(taddr, v);
#pragma unroll 1
            for (int cc = 0; cc < G::NCOLW; cc += 32) {
                const int c0 = cbeg + cc;
                float x[32];
                if (!PREFETCH) tmem_ld32_issue(taddr + (uint32_t)cc, v);    // (four warps per scheduler hide the load; no prefetch registers)
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                {
                    const uint64_t sc2 = pk2(out_scale, out_scale);
#pragma unroll
                    for (int j = 0; j < 32; j += 2) unpk2(fmul2(pk2u(v[j], v[j + 1]), sc2), x[j], x[j + 1]);
                }
                if (cc + 32 < G::NCOLW) { if (PREFETCH) tmem_ld32_issue(taddr + (uint32_t)(cc + 32), v); }          // next chunk in flight
                else {
                    // last chunk of this thread: its TMEM reads are complete -> the accumulator may be overwritten
                    // (tcgen05.wait::ld is warp-collective: all 32 lanes' reads are complete; ONE arrival per warp)
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote_relaxed(mapa_u32(smem_u32(&bars[B_TEMPTY + as]), 0));
                }
                if (P.bias && !(BGNN_EPI_SKIP & 4)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + c0 + j));
                        x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
                    }
                }
                if (P.cmax_out && !(BGNN_EPI_SKIP & 4)) {
                    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        m0 = fmaxf(m0, fabsf(x[j])); m1 = fmaxf(m1, fabsf(x[j + 1]));
                        m2 = fmaxf(m2, fabsf(x[j + 2])); m3 = fmaxf(m3, fabsf(x[j + 3]));
                    }
                    cmax = fmaxf(cmax, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                }
                if (P.alpha && !(BGNN_EPI_SKIP & 1)) {
                    float s0, s1, s2, s3, d0, d1, d2, d3;
                    uint64_t s01 = 0ull, s23 = 0ull, d01 = 0ull, d23 = 0ull;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const ulonglong2 s4 = *reinterpret_cast<const ulonglong2*>(s_par + c0 + j);
                        const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(s_par + PN + c0 + j);
                        const uint64_t x01 = pk2(x[j], x[j + 1]), x23 = pk2(x[j + 2], x[j + 3]);
                        s01 = ffma2(x01, s4.x, s01); s23 = ffma2(x23, s4.y, s23);
                        d01 = ffma2(x01, d4.x, d01); d23 = ffma2(x23, d4.y, d23);
                    }
                    unpk2(s01, s0, s1); unpk2(s23, s2, s3); unpk2(d01, d0, d1); unpk2(d23, d2, d3);
                    as_acc += (s0 + s1) + (s2 + s3);
                    ad_acc += (d0 + d1) + (d2 + d3);
                    if (((c0 + 32) % P.CH) == 0) {          // a head's channels are complete
                        if (row < M) {
                            const int h = (c0 + 32) / P.CH - 1;
                            P.alpha[row * (2 * P.Htot) + h] = as_acc;
                            P.alpha[row * (2 * P.Htot) + P.Htot + h] = ad_acc;
                        }
                        as_acc = 0.f; ad_acc = 0.f;
                    }
                }
                if (BGNN_EPI_SKIP & 2) { if (x[lane] == 123.456f) P.Cout[row] = x[0]; }
                else if constexpr (MLP) {
                    // chunk [32 rows x 32 cols] -> SWIZZLE_128B staging tile -> one TMA tensor store (rows >= M are clipped)
                    unsigned char* tile_s = stg + (size_t)warp * 4096;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // previous store has read the tile
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(tile_s + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const int cx = P.head_major ? (c0 % P.CH) : c0, cz = P.head_major ? (c0 / P.CH) : 0;
                        const int cy = (int)(tile * 2 * TILE_M + rank * TILE_M + quarter * 32);
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];\n"
                                     ::"l"(&tmapC), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(tile_s)) : "memory");
                        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                    }
                } else if (row < M) {           // thread-per-row 256-bit stores
                    float* dst = P.head_major ? P.Cout + ((int64_t)(c0 / P.CH) * M + row) * P.CH + (c0 % P.CH) : P.Cout + row * PN + c0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) st_global_v8(dst + 8 * j, x + 8 * j);
                }
            }
#ifdef BGNN_AGG_TIMING
            if (warp == 0 && lane == 0 && rank == 0) UT_TICKP(4);
#endif
        }
        if (MLP && lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");     // all tensor stores done
        if (P.cmax_out) {
#pragma unroll
            for (int o = 16; o; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
            if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(P.cmax_out), __float_as_uint(cmax));
        }
    }
    // teardown: both CTAs are done with TMEM and with each other's barriers before either frees / leaves
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == MMA_W) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D tensor map over A [M, K] fp32 (row-major) with a [box_rows x 32] box; out-of-range rows read as zero
int get_encode_fn(EncodeTiledFn* out) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        BGNN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { bgnn_set_error("cuTensorMapEncodeTiled is not available from this driver"); return BGNN_E_CUDA; }
        fn = (EncodeTiledFn)p;
    }
    *out = fn;
    return BGNN_OK;
}

int make_tmap_A(const float* A, int64_t M, int K, int box_rows, CUtensorMap* out) {
    EncodeTiledFn fn = nullptr;
    { int rc = get_encode_fn(&fn); if (rc) return rc; }
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)K * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { bgnn_set_error("cuTensorMapEncodeTiled failed (%d) for A=%p M=%lld K=%d", (int)r, (const void*)A, (long long)M, K); return BGNN_E_CUDA; }
    return BGNN_OK;
}

// 2-D tensor map over one fp16 plane of a pre-split A [M, K] with a [box_rows x 32] SWIZZLE_64B box (64 B inner extent):
// the box lands as 128-row K-major swizzled tiles, rows past M read as zero
int make_tmap_A16(const uint16_t* A, int64_t M, int K, int64_t ld, int box_rows, CUtensorMap* out) {
    EncodeTiledFn fn = nullptr;
    { int rc = get_encode_fn(&fn); if (rc) return rc; }
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(uint16_t)};
    const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(A), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { bgnn_set_error("cuTensorMapEncodeTiled failed (%d) for split A=%p M=%lld K=%d", (int)r, (const void*)A, (long long)M, K); return BGNN_E_CUDA; }
    return BGNN_OK;
}

// the same plane with a [box_rows x 64] SWIZZLE_128B box (128 B inner extent; K % 64 == 0): proj_pair_kernel's A units
int make_tmap_A16_wide(const uint16_t* A, int64_t M, int K, int64_t ld, int box_rows, CUtensorMap* out) {
    EncodeTiledFn fn = nullptr;
    { int rc = get_encode_fn(&fn); if (rc) return rc; }
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(uint16_t)};
    const cuuint32_t box[2] = {(cuuint32_t)pairk::KCA, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(A), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { bgnn_set_error("cuTensorMapEncodeTiled failed (%d) for split A=%p M=%lld K=%d (128 B box)", (int)r, (const void*)A, (long long)M, K); return BGNN_E_CUDA; }
    return BGNN_OK;
}

// 3-D tensor map over the projection output with a [32 cols x 32 rows x 1] SWIZZLE_128B box:
//   head-major  C[Htot][M][CH]: dims {CH, M, Htot}      row-major  C[M][N]: dims {N, M, 1}
int make_tmap_C(float* C, int64_t M, int N, int CH, int head_major, CUtensorMap* out) {
    EncodeTiledFn fn = nullptr;
    int rc = get_encode_fn(&fn);
    if (rc) return rc;
    const int Htot = head_major ? N / CH : 1;
    const int inner = head_major ? CH : N;
    const cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)M, (cuuint64_t)Htot};
    const cuuint64_t gstr[2] = {(cuuint64_t)inner * sizeof(float), (cuuint64_t)M * inner * sizeof(float)};
    const cuuint32_t box[3] = {32, 32, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, C, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { bgnn_set_error("cuTensorMapEncodeTiled failed (%d) for C=%p M=%lld N=%d", (int)r, (void*)C, (long long)M, N); return BGNN_E_CUDA; }
    return BGNN_OK;
}

template <int N, int PMODE, int EMODE, int NH = 1>
int launch_t(const UmmaArgs& a, int sms, cudaStream_t s) {
    using S = UmmaSmem<N, PMODE, EMODE, NH>;
    static_assert(S::BYTES <= 232448, "shared memory budget exceeded");
    auto kern = proj_umma_kernel<N, PMODE, EMODE, NH>;
    BGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES));   // per device, idempotent, thread-safe
    alignas(64) CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (PMODE == PM_PLAIN) {
        BGNN_CHECK_ARG((reinterpret_cast<uintptr_t>(a.A) & 15) == 0, "A must be 16-byte aligned");
        int rc = make_tmap_A(a.A, a.M, a.K, S::SUPER_M, &tmap);
        if (rc) return rc;
    }
    alignas(64) CUtensorMap tmap2;
    memset(&tmap2, 0, sizeof(tmap2));
    if (PMODE == PM_SPLIT) {
        BGNN_CHECK_ARG(a.A_hi && a.A_lo && ((reinterpret_cast<uintptr_t>(a.A_hi) | reinterpret_cast<uintptr_t>(a.A_lo)) & 15) == 0,
                       "split A planes must be 16-byte aligned");
        const int64_t ld = a.a_ld > 0 ? a.a_ld : a.K;
        int rc = make_tmap_A16(a.A_hi, a.M, a.K, ld, S::SUPER_M, &tmap);
        if (!rc) rc = make_tmap_A16(a.A_lo, a.M, a.K, ld, S::SUPER_M, &tmap2);
        if (rc) return rc;
    }
    alignas(64) CUtensorMap tmapC;
    memset(&tmapC, 0, sizeof(tmapC));
    if (EMODE == EM_PROJ) {
        BGNN_CHECK_ARG((reinterpret_cast<uintptr_t>(a.Cout) & 15) == 0, "C must be 16-byte aligned");
        int rc = make_tmap_C(a.Cout, a.M, S::NT, a.CH, a.head_major, &tmapC);
        if (rc) return rc;
    }
    const int64_t n_super = (a.M + S::SUPER_M - 1) / S::SUPER_M;
    unsigned grid = (unsigned)(n_super < sms ? n_super : sms);
    if (NH > 1) {                                   // pairs of CTAs (one per column block) share the M tiles
        const int64_t pairs = n_super < sms / 2 ? n_super : sms / 2;
        grid = (unsigned)(2 * (pairs > 0 ? pairs : 1));
    }
    static const bool mc_env = getenv("BGNN_NO_MULTICAST") == nullptr;        // on by default (A/B: tools/ab_mc.sh)
    if (PMODE == PM_SPLIT && NH == 2 && mc_env) {
        UmmaArgs am = a;
        am.mc = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(UM_THREADS); cfg.dynamicSmemBytes = S::BYTES; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const cudaError_t el = cudaLaunchKernelEx(&cfg, kern, am, tmap, tmap2, tmapC);
        if (el != cudaSuccess) { bgnn_set_error("cluster launch of proj_umma_kernel failed: %s", cudaGetErrorString(el)); return BGNN_E_CUDA; }
        return BGNN_OK;
    }
    kern<<<grid, UM_THREADS, S::BYTES, s>>>(a, tmap, tmap2, tmapC);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        bgnn_set_error("proj_umma_kernel<N=%d x %d, producer %d, epilogue %d> M=%lld K=%d: %s", N, NH, PMODE, EMODE, (long long)a.M, a.K,
                       cudaGetErrorString(e));
        return BGNN_E_CUDA;
    }
    return BGNN_OK;
}

// cta_group::2 pair kernel (EM_PROJ, N = 256): PM_SPLIT (K <= 256, K % 64 == 0) or PM_MLP0 (K == 64)
template <int MODE>
int launch_pair(const UmmaArgs& a, int sms, cudaStream_t s) {
    using G = pairk::Cfg<MODE>;
    auto kern = proj_pair_kernel<MODE>;
    BGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::BYTES));
    alignas(64) CUtensorMap tmap, tmap2, tmapC;
    memset(&tmap, 0, sizeof(tmap)); memset(&tmap2, 0, sizeof(tmap2)); memset(&tmapC, 0, sizeof(tmapC));
    BGNN_CHECK_ARG((reinterpret_cast<uintptr_t>(a.Cout) & 31) == 0, "C must be 32-byte aligned");
    if (MODE == PM_SPLIT) {
        BGNN_CHECK_ARG(a.A_hi && a.A_lo && ((reinterpret_cast<uintptr_t>(a.A_hi) | reinterpret_cast<uintptr_t>(a.A_lo)) & 15) == 0,
                       "split A planes must be 16-byte aligned");
        const int64_t ld = a.a_ld > 0 ? a.a_ld : a.K;
        int rc = make_tmap_A16_wide(a.A_hi, a.M, a.K, ld, TILE_M, &tmap);
        if (!rc) rc = make_tmap_A16_wide(a.A_lo, a.M, a.K, ld, TILE_M, &tmap2);
        if (rc) return rc;
    } else {
        int rc = make_tmap_C(a.Cout, a.M, pairk::PN, a.CH, a.head_major, &tmapC);
        if (rc) return rc;
    }
    const int64_t n_tiles = (a.M + 2 * TILE_M - 1) / (2 * TILE_M);
    const int64_t pairs = n_tiles < sms / 2 ? n_tiles : sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * (pairs > 0 ? pairs : 1))); cfg.blockDim = dim3(G::THREADS); cfg.dynamicSmemBytes = G::BYTES; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const cudaError_t el = cudaLaunchKernelEx(&cfg, kern, a, tmap, tmap2, tmapC);
    if (el != cudaSuccess) { bgnn_set_error("cluster launch of proj_pair_kernel<%d> failed: %s", MODE, cudaGetErrorString(el)); return BGNN_E_CUDA; }
    return BGNN_OK;
}

// max |v| over n floats -> atomicMax on the bit pattern (non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
    float m = 0.f;
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(v) & 15) == 0) ? n / 4 : 0;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = gid; i < n4; i += gsz) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(v) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(q.x), fabsf(q.y))), fmaxf(fabsf(q.z), fabsf(q.w)));
    }
    for (int64_t i = n4 * 4 + gid; i < n; i += gsz) m = fmaxf(m, fabsf(__ldg(v + i)));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}

// host fp16 conversions (round to nearest even)
uint16_t f32_to_f16_rn(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | (x > 0x7f800000u ? 0x7e00u : 0x7c00u));
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                 // >= 65520 rounds to inf
    if (x < 0x33000001u) return sign;                                        // <= 2^-25 rounds to zero
    if (x < 0x38800000u) {                                                   // subnormal half (< 2^-14): units of 2^-24
        const int e = (int)(x >> 23);
        const uint32_t mant = (x & 0x7fffffu) | 0x800000u;
        const int shift = 126 - e;
        uint32_t r = mant >> shift;
        const uint32_t rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (r & 1u))) ++r;
        return (uint16_t)(sign | r);
    }
    const uint32_t e = (x >> 23) - 112u, mant = x & 0x7fffffu;
    uint32_t r = (e << 10) | (mant >> 13);
    const uint32_t rem = mant & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;
    return (uint16_t)(sign | r);
}
float f16_to_f32(uint16_t h) {
    const int e = (h >> 10) & 31, m = h & 1023;
    float v;
    if (e == 0) v = ldexpf((float)m, -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = ldexpf((float)(m | 1024), e - 25);
    return (h & 0x8000) ? -v : v;
}

static size_t sw64_off_host(int r, int c) { return (size_t)(r >> 3) * SBO_BYTES + (r & 7) * 64 + ((c ^ ((r & 7) >> 1)) << 4); }

}  // namespace

bool proj_umma_supported(int N, int K, int CH, bool want_alpha) {
    if (!(N == 32 || N == 64 || N == 128 || N == 256)) return false;
    if (K < KC || K % KC != 0) return false;
    if (want_alpha && (CH % 32 != 0 || N % CH != 0)) return false;
    if (want_alpha && N >= 256 && (N / 2) % CH != 0) return false;      // column halves must hold whole heads
    return true;
}

// host: W [N,K] fp32 -> fp16 splits of W*s (s = power of two with max|W*s| < 1): hi = fp16(W s), lo = fp16((W s - hi) 2^11),
// arranged as [K/32][2][N x 32 tile] where each tile is ALREADY in the K-major SWIZZLE_64B shared-memory layout the MMA
// descriptors expect (one bulk copy per K chunk).  Returns 1/s.
float proj_umma_split_weights(const float* W, int N, int K, uint16_t* out /*[K/32][2][N*32]*/) {
    float mx = 0.f;
    for (size_t i = 0; i < (size_t)N * K; ++i) mx = fmaxf(mx, fabsf(W[i]));
    int e = 0;
    if (mx > 0.f && isfinite(mx)) frexpf(mx, &e);          // mx = m * 2^e, m in [0.5, 1)
    const float s = ldexpf(1.f, -e), inv = ldexpf(1.f, e);
    const size_t tile = (size_t)N * KC;
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const float w = W[(size_t)n * K + k] * s;
            const uint16_t hi = f32_to_f16_rn(w);
            const uint16_t lo = f32_to_f16_rn((w - f16_to_f32(hi)) * LO_SCALE);
            const int kc = k / KC, kk = k % KC;
            const size_t el = sw64_off_host(n, kk >> 3) / 2 + (kk & 7);
            uint16_t* base = out + (size_t)kc * NSPLIT * tile;
            base[el] = hi; base[tile + el] = lo;
        }
    return inv;
}

bool proj_umma_presplit_supported(int N, int K, int CH, bool want_alpha) {
    return (N == 64 || N == 256) && proj_umma_supported(N, K, CH, want_alpha);
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

bool proj_umma_heads_presplit_supported(int K, int Npad) { return Npad == 128 && K % KC == 0 && K >= KC; }

int launch_absmax(const float* v, int64_t n, float* out, int sms, cudaStream_t s) {
    if (n <= 0) return BGNN_OK;
    int64_t blocks = ceil_div64(ceil_div64(n, 4), 256);
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    absmax_kernel<<<(unsigned)blocks, 256, 0, s>>>(v, n, out);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

static const bool g_mlp0_whole = getenv("BGNN_MLP0_WHOLE") != nullptr;       // dev switch: one 128 x 256 tile per CTA (1.30 vs 1.25 ms paired)
static const bool g_pair_mlp0 = getenv("BGNN_NO_CG2_MLP0") == nullptr && getenv("BGNN_NO_CG2") == nullptr;   // layer-0 projection (fused MLP) as a
                                                                         // cta_group::2 pair: on by default (1.03 vs 1.38 ms)
static const bool g_pair_cg2 = getenv("BGNN_NO_CG2") == nullptr;             // on by default; dev switch back to the N = 2 x 128 multicast pair
static const bool g_split_whole = getenv("BGNN_SPLIT_WHOLE") != nullptr;     // dev switch: pre-split A, N = 256 as one 128 x 256 tile per CTA
static const bool g_n256_halves = getenv("BGNN_N256_HALVES") != nullptr;     // dev switch (measured slower with fp32 A: 2.49 vs 2.13 ms; the
                                                                         // 2-deep operand ring cannot hide the split latency)

int launch_umma(const UmmaArgs& a, int N, int pmode, int emode, int sms, cudaStream_t s) {
    if (a.M == 0) return BGNN_OK;
    if (pmode == PM_PLAIN && emode == EM_PROJ) {
        switch (N) {
            case 32: return launch_t<32, PM_PLAIN, EM_PROJ>(a, sms, s);
            case 64: return launch_t<64, PM_PLAIN, EM_PROJ>(a, sms, s);
            case 128: return launch_t<128, PM_PLAIN, EM_PROJ>(a, sms, s);
            case 256:
                // paired CTAs, one column block of 128 each, resident W, double-buffered accumulators; an epilogue warp then
                // owns 64 columns, which must hold whole heads
                if (g_n256_halves && a.K <= 8 * KC && (!a.alpha || 64 % a.CH == 0)) return launch_t<128, PM_PLAIN, EM_PROJ, 2>(a, sms, s);
                return launch_t<256, PM_PLAIN, EM_PROJ>(a, sms, s);
        }
    } else if (pmode == PM_SPLIT && emode == EM_PROJ) {
        switch (N) {
            case 64: return launch_t<64, PM_SPLIT, EM_PROJ>(a, sms, s);
            case 256:
                if (g_pair_cg2 && a.K <= 8 * KC && a.K % pairk::KCA == 0 && (!a.alpha || pairk::Cfg<PM_SPLIT>::NCOLW % a.CH == 0)) return launch_pair<PM_SPLIT>(a, sms, s);
                if (!g_split_whole && a.K <= 8 * KC && (!a.alpha || 64 % a.CH == 0)) return launch_t<128, PM_SPLIT, EM_PROJ, 2>(a, sms, s);
                return launch_t<256, PM_SPLIT, EM_PROJ>(a, sms, s);
        }
    } else if (pmode == PM_MLP0 && emode == EM_PROJ) {
        switch (N) {
            case 128: return launch_t<128, PM_MLP0, EM_PROJ>(a, sms, s);
            case 256:
                // paired CTAs, N = 2 x 128 (TMA stores): the epilogue, which dominates at K = 64, overlaps the next tile's MLP + MMAs
                if (g_pair_mlp0 && a.K == pairk::KCA && a.F <= MLP0_FP && (!a.alpha || pairk::Cfg<PM_MLP0>::NCOLW % a.CH == 0)) return launch_pair<PM_MLP0>(a, sms, s);
                if (!g_mlp0_whole && a.K <= 8 * KC && (!a.alpha || 64 % a.CH == 0)) return launch_t<128, PM_MLP0, EM_PROJ, 2>(a, sms, s);
                return launch_t<256, PM_MLP0, EM_PROJ>(a, sms, s);
        }
    } else if (pmode == PM_SPLIT && emode == EM_HEADS) {
        switch (N) {
            case 128: return launch_t<128, PM_SPLIT, EM_HEADS>(a, sms, s);
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

// fp32 rows -> fp16 hi / lo planes of a * sA (test aid: what the tiled aggregation writes for PM_SPLIT)
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ A, int64_t n, const float* __restrict__ amax,
                                                           uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const float bound = __ldg(amax);
    const int e = (int)((__float_as_uint(bound) >> 23) & 0xffu), es = 253 - e;
    const float sA = (e > 0 && e < 255 && es >= 1 && es <= 253) ? __uint_as_float((uint32_t)es << 23) : 1.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = __ldg(A + i) * sA;
        const __half h = __float2half_rn(a);
        const __half l = __float2half_rn((a - __half2float(h)) * LO_SCALE);
        hi[i] = *reinterpret_cast<const uint16_t*>(&h);
        lo[i] = *reinterpret_cast<const uint16_t*>(&l);
    }
}

static int test_proj_umma_impl(const float* A_dev, int64_t M, int K, const float* W_host, int N, float* C_dev,
                               const float* att_src_dev, const float* att_dst_dev, float* alpha_dev, int heads,
                               void* stream_, bool presplit);

// ---- test hooks (declared in include/bgnn.h): C = A . W^T on the tensor-core path, W given as fp32 on the HOST
extern "C" int bgnn_test_proj_umma(const float* A_dev, int64_t M, int K, const float* W_host, int N, float* C_dev,
                                   const float* att_src_dev, const float* att_dst_dev, float* alpha_dev, int heads,
                                   void* stream_) {
    return test_proj_umma_impl(A_dev, M, K, W_host, N, C_dev, att_src_dev, att_dst_dev, alpha_dev, heads, stream_, false);
}
extern "C" int bgnn_test_proj_umma_presplit(const float* A_dev, int64_t M, int K, const float* W_host, int N, float* C_dev,
                                            const float* att_src_dev, const float* att_dst_dev, float* alpha_dev, int heads,
                                            void* stream_) {
    return test_proj_umma_impl(A_dev, M, K, W_host, N, C_dev, att_src_dev, att_dst_dev, alpha_dev, heads, stream_, true);
}

static int test_proj_umma_impl(const float* A_dev, int64_t M, int K, const float* W_host, int N, float* C_dev,
                               const float* att_src_dev, const float* att_dst_dev, float* alpha_dev, int heads,
                               void* stream_, bool presplit) {
    BGNN_CHECK_ARG(A_dev && W_host && C_dev, "null argument");
    const int CH = heads > 0 ? N / heads : N;
    BGNN_CHECK_ARG(proj_umma_supported(N, K, CH, alpha_dev != nullptr), "shape N=%d K=%d not supported by the tcgen05 path", N, K);
    BGNN_CHECK_ARG(!presplit || proj_umma_presplit_supported(N, K, CH, alpha_dev != nullptr), "shape N=%d K=%d has no pre-split variant", N, K);
    cudaStream_t s = (cudaStream_t)stream_;
    std::vector<uint16_t> split((size_t)NSPLIT * N * K);
    const float winv = proj_umma_split_weights(W_host, N, K, split.data());
    uint16_t* d = nullptr;
    float* amax = nullptr;
    BGNN_CUDA(cudaMalloc(&d, split.size() * 2));
    BGNN_CUDA(cudaMalloc(&amax, 2 * sizeof(float)));
    BGNN_CUDA(cudaMemsetAsync(amax, 0, 2 * sizeof(float), s));
    BGNN_CUDA(cudaMemcpyAsync(d, split.data(), split.size() * 2, cudaMemcpyHostToDevice, s));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int rc = launch_absmax(A_dev, M * K, amax, sms, s);
    if (rc == BGNN_OK) {
        UmmaArgs a = {};
        a.A = A_dev; a.M = M; a.K = K; a.Wsplit = d; a.w_inv_scale = winv;
        a.amax_src = amax; a.amax_mul = 1.f; a.amax_add = 0.f; a.cmax_out = amax + 1;
        a.Cout = C_dev; a.att_src = att_src_dev; a.att_dst = att_dst_dev; a.alpha = alpha_dev;
        a.Htot = alpha_dev ? heads : 1; a.CH = alpha_dev ? CH : N;
        uint16_t* planes = nullptr;
        if (presplit) {
            BGNN_CUDA(cudaMalloc(&planes, (size_t)M * K * 4));
            split_planes_kernel<<<sms * 8, 256, 0, s>>>(A_dev, M * K, amax, planes, planes + (size_t)M * K);
            a.A_hi = planes; a.A_lo = planes + (size_t)M * K;
        }
        rc = launch_umma(a, N, presplit ? PM_SPLIT : PM_PLAIN, EM_PROJ, sms, s);
        cudaStreamSynchronize(s);
        if (planes) cudaFree(planes);
    }
    cudaStreamSynchronize(s);
    cudaFree(d);
    cudaFree(amax);
    return rc;
}
