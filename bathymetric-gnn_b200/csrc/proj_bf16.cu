// bf16-mode projection GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//     C[M, N] = A[M, K] . W[N, K]^T        A, W, C bf16 in HBM, fp32 accumulation in TMEM
//
// The `h = x W^T` of every GATConv layer (PyG GATConv.lin, called from models/gnn.py:175-176) in the bf16 mode of
// bgnn_gat_forward (include/bgnn.h, BGNN_PREC_BF16; BASELINE configs[4]: k = 24, hidden 128, H*C = 512).  Both operands are
// bf16 already (the previous aggregation wrote A as bf16 rows, the weights were rounded once), so -- unlike the fp32-parity
// kernel in proj_umma.cu, which needs a 3-product fp16 split -- ONE tcgen05.mma.kind::f16 product per K step does the job
// and no conversion stage exists: TMA -> shared memory -> tensor core.
//
// Persistent CTA per SM, 320 threads, warp-specialised:
//   (320 threads since the epilogue runs as two groups of four warps, see below)
//   warp 0 (one thread)  TMA producer.  The CTA owns ONE 128-column block of W for the whole launch: its [BN x K] slab
//                        (<= 128 KB) is fetched once and stays resident in shared memory (re-reading W per 128-row tile
//                        would double the L2 -> SM traffic and make the kernel L2-bound).  A streams through a 3-6 deep ring
//                        of [128 x 64] SWIZZLE_128B tiles, one cp.async.bulk.tensor per K chunk (rows past M read as zero).
//   warp 1 (one thread)  MMA issuer: 4 x tcgen05.mma (M = 128, N = BN, K = 16) per chunk; tcgen05.commit frees the A stage;
//                        accumulators double-buffered in TMEM (2 x BN columns) so the epilogue of tile t overlaps tile t+1.
//   warps 2-9            epilogue, two groups of four warps: group g takes the tiles of accumulator stage g, so two tiles are
//                        finished at a time (one warp per scheduler runs ~2 000 dependent instructions per tile: a single
//                        group was the critical path, ncu: tensor pipe 27 % busy, issue slots 28 %).  Per tile: tcgen05.ld 32x32b.x32 (thread <-> row), alpha_src / alpha_dst dot products per head from
//                        the fp32 accumulators, RN conversion to bf16, rows staged in shared memory and written out
//                        coalesced (16 B per lane, consecutive lanes <-> consecutive pieces of a row) into the head-major
//                        layout [Htot][M][CH] the aggregation reads.
// The column blocks of one 128-row tile are handled by CTAs b, b+1, .. (same tile sequence), so A comes from HBM once and from
// L2 for the other blocks.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"
#include "proj_umma.cuh"

namespace {

constexpr int KC = 64;                       // K elements per stage: 128 B of bf16 = one SWIZZLE_128B row
constexpr int TILE_M = 128;
constexpr int A_TILE_BYTES = TILE_M * KC * 2;        // 16 KB
constexpr int MAX_STAGE = 6;                 // A ring depth: as many 16 KB stages as fit beside the resident W slab (>= 3)
constexpr uint32_t SBO_BYTES = 8 * KC * 2;           // 1024 B between 8-row groups
constexpr int EPI_WARPS = 8;                 // two groups of four: group g finishes the tiles whose accumulator stage is g
constexpr int THREADS = (2 + EPI_WARPS) * 32;        // 320
constexpr int MAX_K = 512;

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
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp bit layout; layout type 2 = SWIZZLE_128B).
// The fp32-parity kernel uses SWIZZLE_64B tiles (KC = 32); 128-byte rows halve the number of TMA copies / barrier round trips
// per tile and read shared memory without the 2-way conflicts of 64-byte rows.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((SBO_BYTES >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
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
__device__ __forceinline__ uint32_t pack2(float a, float b) {      // one packed convert (F2FP.BF16.F32.PACK_AB), a in the low half
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
}

struct Bf16Args {
    int64_t M; int K; int N; int NB;          // NB = N / BN column blocks
    uint16_t* C; const float* att_src; const float* att_dst; float* alpha; int Htot; int CH; int ch_shift;   // CH = 1 << ch_shift
    int nstage;
};

template <int BN> struct Smem {
    static constexpr int W_CHUNK_BYTES = BN * KC * 2;                 // one K chunk of the resident W slab
    static constexpr int PITCH = BN + 16;                              // staging row pitch (bytes) for HALF the columns of a tile:
                                                                       // conflict-free 16 B accesses; staging the full width
                                                                       // would leave only 3 A stages at K = 512 (measured: slower)
    static constexpr int STAGE_ROWS_BYTES = 32 * PITCH;                // per epilogue warp
    static constexpr size_t fixed(int K) {
        return 1024 /*alignment slack*/ + (size_t)(K / KC) * W_CHUNK_BYTES + (size_t)EPI_WARPS * STAGE_ROWS_BYTES +
               2 * BN * sizeof(float) + 64 * sizeof(uint64_t);
    }
    static int stages(int K) {
        const long room = 232448 - 1024 /*static + reserved*/ - (long)fixed(K);
        const int n = (int)(room / A_TILE_BYTES);
        return n > MAX_STAGE ? MAX_STAGE : n;
    }
    static size_t bytes(int K) { return fixed(K) + (size_t)stages(K) * A_TILE_BYTES; }
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
proj_bf16_kernel(const __grid_constant__ Bf16Args P, const __grid_constant__ CUtensorMap tmapA,
                 const __grid_constant__ CUtensorMap tmapW) {
    using S = Smem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nkc = P.K / KC;
    uint8_t* wres = base;                                               // [nkc][BN x 64 B] swizzled
    uint8_t* aring = wres + (size_t)nkc * S::W_CHUNK_BYTES;             // [NSTAGE][128 x 64 B] swizzled
    const int NSTAGE = P.nstage;
    uint8_t* stg = aring + (size_t)NSTAGE * A_TILE_BYTES;               // [EPI_WARPS][32][PITCH]
    float* s_att = reinterpret_cast<float*>(stg + (size_t)EPI_WARPS * S::STAGE_ROWS_BYTES);     // att_src [BN] | att_dst [BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_att + 2 * BN);
    uint64_t* bar_full = bars;                 // [MAX_STAGE]
    uint64_t* bar_empty = bars + MAX_STAGE;    // [MAX_STAGE]
    uint64_t* bar_w = bars + 2 * MAX_STAGE;    // resident W landed
    uint64_t* bar_tfull = bar_w + 1;           // [2]
    uint64_t* bar_tempty = bar_tfull + 2;      // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = blockIdx.x % P.NB;                      // this CTA's column block
    const int grp = blockIdx.x / P.NB, ngrp = gridDim.x / P.NB;
    const int64_t n_mt = (P.M + TILE_M - 1) / TILE_M;
    const int n0 = nb * BN;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(bar_w, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 4 * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < BN; i += THREADS) {
        s_att[i] = P.alpha ? __ldg(P.att_src + n0 + i) : 0.f;
        s_att[BN + i] = P.alpha ? __ldg(P.att_dst + n0 + i) : 0.f;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            {   // resident W slab: nkc boxes of [BN rows x 32 k]
                const uint32_t bar = smem_u32(bar_w);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)(nkc * S::W_CHUNK_BYTES)) : "memory");
                for (int kc = 0; kc < nkc; ++kc)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                                 ::"r"(smem_u32(wres + (size_t)kc * S::W_CHUNK_BYTES)), "l"(&tmapW), "r"(kc * KC), "r"(n0), "r"(bar) : "memory");
            }
            uint32_t it = 0;
            for (int64_t mt = grp; mt < n_mt; mt += ngrp) {
                const int row0 = (int)(mt * TILE_M);
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int s = it % NSTAGE;
                    if (it >= (uint32_t)NSTAGE) mbar_wait(&bar_empty[s], ((it / NSTAGE) - 1) & 1);
                    const uint32_t bar = smem_u32(&bar_full[s]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)A_TILE_BYTES) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                                 ::"r"(smem_u32(aring + (size_t)s * A_TILE_BYTES)), "l"(&tmapA), "r"(kc * KC), "r"(row0), "r"(bar) : "memory");
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            // instruction descriptor: D = F32 (bit 4), A = B = BF16 (format 1 at bits 7 and 10), both K-major, N, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            mbar_wait(bar_w, 0);
            uint32_t it = 0, t = 0;
            for (int64_t mt = grp; mt < n_mt; mt += ngrp, ++t) {
                const uint32_t as = t & 1;
                if (t >= 2) mbar_wait(&bar_tempty[as], ((t >> 1) - 1) & 1);        // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t d = tmem_base + as * BN;
                for (int kc = 0; kc < nkc; ++kc, ++it) {
                    const int s = it % NSTAGE;
                    mbar_wait(&bar_full[s], (it / NSTAGE) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t sA = smem_u32(aring + (size_t)s * A_TILE_BYTES), sB = smem_u32(wres + (size_t)kc * S::W_CHUNK_BYTES);
#pragma unroll
                    for (int ks = 0; ks < KC / 16; ++ks)
                        umma_bf16(d, make_desc(sA + ks * 32), make_desc(sB + ks * 32), idesc, (kc == 0 && ks == 0) ? 0u : 1u);
                    umma_commit(&bar_empty[s]);
                }
                umma_commit(&bar_tfull[as]);
            }
        }
    } else {
        // ===================== epilogue: thread <-> row of the tile =====================
        const int ew = warp - 2;                         // staging slot
        const int lg = warp & 3;                         // TMEM lane group this warp may access
        const uint32_t egrp = (uint32_t)ew >> 2;         // tiles with accumulator stage == egrp are this warp's
        uint8_t* my_stg = stg + (size_t)ew * S::STAGE_ROWS_BYTES;
        const int CH = P.CH, Htot = P.Htot, chs = P.ch_shift;
        uint32_t t = 0;
        for (int64_t mt = grp; mt < n_mt; mt += ngrp, ++t) {
            const uint32_t as = t & 1;
            if (as != egrp) continue;
            mbar_wait(&bar_tfull[as], (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int64_t row = mt * TILE_M + lg * 32 + lane;
            float ps = 0.f, pd = 0.f;
            constexpr int CPH = BN / 64;                       // 32-column chunks per half tile
            constexpr int PPR = BN / 16;                       // 16-byte pieces per staged (half) row
            const int64_t rbase = mt * TILE_M + lg * 32;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll 1
                for (int cc = 0; cc < CPH; ++cc) {
                    const int c = hf * CPH + cc;
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + as * BN + c * 32;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                    if (P.alpha) {
                        // four partial sums per dot product: the 32-long dependent FMA chains were a third of the chunk's latency
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 a4 = *reinterpret_cast<const float4*>(s_att + c * 32 + j);
                            const float4 b4 = *reinterpret_cast<const float4*>(s_att + BN + c * 32 + j);
                            const float x0 = __uint_as_float(v[j]), x1 = __uint_as_float(v[j + 1]);
                            const float x2 = __uint_as_float(v[j + 2]), x3 = __uint_as_float(v[j + 3]);
                            s0 = fmaf(x0, a4.x, s0); s1 = fmaf(x1, a4.y, s1); s2 = fmaf(x2, a4.z, s2); s3 = fmaf(x3, a4.w, s3);
                            d0 = fmaf(x0, b4.x, d0); d1 = fmaf(x1, b4.y, d1); d2 = fmaf(x2, b4.z, d2); d3 = fmaf(x3, b4.w, d3);
                        }
                        ps += (s0 + s1) + (s2 + s3);
                        pd += (d0 + d1) + (d2 + d3);
                        if ((((c + 1) * 32) & (CH - 1)) == 0) {   // a head is complete (CH is a power of two >= 32)
                            if (row < P.M) {
                                const int head = (n0 + c * 32) >> chs;
                                P.alpha[row * (2 * Htot) + head] = ps;
                                P.alpha[row * (2 * Htot) + Htot + head] = pd;
                            }
                            ps = 0.f; pd = 0.f;
                        }
                    }
                    uint4* dst = reinterpret_cast<uint4*>(my_stg + (size_t)lane * S::PITCH + cc * 64);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[q] = make_uint4(pack2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1])),
                                            pack2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])),
                                            pack2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])),
                                            pack2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
                }
                if (hf == 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                    mbar_arrive(&bar_tempty[as]);              // every accumulator column of this row has been read
                }
                __syncwarp();
                // coalesced write-out of this half: 16 B pieces, consecutive lanes <-> consecutive pieces of a row
#pragma unroll 4
                for (int p = lane; p < 32 * PPR; p += 32) {
                    const int r = p / PPR, pc = p % PPR;
                    const int64_t m = rbase + r;
                    if (m < P.M) {
                        const int n = n0 + hf * (BN / 2) + pc * 8;
                        const uint4 val = *reinterpret_cast<const uint4*>(my_stg + (size_t)r * S::PITCH + pc * 16);
                        *reinterpret_cast<uint4*>(P.C + ((((int64_t)(n >> chs) * P.M + m) << chs) + (n & (CH - 1)))) = val;
                    }
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_fn(EncodeTiledFn* out) {
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

// 2-D tensor map over a row-major bf16 matrix [rows, K] with a [box_rows x 64] SWIZZLE_128B box; rows past the end read as zero
int make_tmap_bf16(const uint16_t* base, int64_t rows, int K, int box_rows, CUtensorMap* out) {
    EncodeTiledFn fn = nullptr;
    { int rc = encode_fn(&fn); if (rc) return rc; }
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)K * sizeof(uint16_t)};
    const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { bgnn_set_error("cuTensorMapEncodeTiled failed (%d) for bf16 matrix %p rows=%lld K=%d", (int)r, (const void*)base, (long long)rows, K); return BGNN_E_CUDA; }
    return BGNN_OK;
}

template <int BN>
int launch_t(const Bf16Args& a, const uint16_t* A, const uint16_t* W, int sms, cudaStream_t s) {
    alignas(64) CUtensorMap tA, tW;
    int rc;
    if ((rc = make_tmap_bf16(A, a.M, a.K, TILE_M, &tA))) return rc;
    if ((rc = make_tmap_bf16(W, a.N, a.K, BN, &tW))) return rc;
    const size_t smem = Smem<BN>::bytes(a.K);
    BGNN_CUDA(cudaFuncSetAttribute(proj_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<BN>::bytes(a.K)));
    const int64_t n_mt = (a.M + TILE_M - 1) / TILE_M;
    int64_t groups = sms / a.NB;
    if (groups < 1) groups = 1;
    if (groups > n_mt) groups = n_mt;
    proj_bf16_kernel<BN><<<(unsigned)(groups * a.NB), THREADS, smem, s>>>(a, tA, tW);
    BGNN_LAUNCH_CHECK();
    return BGNN_OK;
}

}  // namespace

bool proj_bf16_supported(int N, int K, int CH) {
    if (!(N == 64 || N == 128 || N == 256 || N == 512)) return false;
    const int BN = N >= 128 ? 128 : 64;
    if (K < KC || K % KC != 0 || K > MAX_K) return false;
    if (!(CH == 32 || CH == 64 || CH == 128) || BN % CH != 0) return false;
    return true;
}

int launch_proj_bf16(const uint16_t* A, const uint16_t* W, int64_t M, int K, int N, uint16_t* Cout, const float* att_src,
                     const float* att_dst, float* alpha, int Htot, int CH, int sms, cudaStream_t s) {
    if (M == 0) return BGNN_OK;
    if (!proj_bf16_supported(N, K, CH)) { bgnn_set_error("bf16 projection N=%d K=%d CH=%d is not compiled", N, K, CH); return BGNN_E_UNSUPPORTED; }
    if (M >= (int64_t)1 << 31) { bgnn_set_error("too many rows for the tensor map"); return BGNN_E_UNSUPPORTED; }
    const int BN = N >= 128 ? 128 : 64;
    Bf16Args a = {};
    a.M = M; a.K = K; a.N = N; a.NB = N / BN; a.C = Cout; a.att_src = att_src; a.att_dst = att_dst; a.alpha = alpha;
    a.Htot = Htot; a.CH = CH; a.ch_shift = CH == 32 ? 5 : CH == 64 ? 6 : 7;
    a.nstage = BN == 128 ? Smem<128>::stages(K) : Smem<64>::stages(K);
    return BN == 128 ? launch_t<128>(a, A, W, sms, s) : launch_t<64>(a, A, W, sms, s);
}

extern "C" int bgnn_test_proj_bf16(const uint16_t* A_dev, int64_t M, int K, const uint16_t* W_dev, int N, uint16_t* C_dev,
                                   const float* att_src_dev, const float* att_dst_dev, float* alpha_dev, int heads,
                                   void* stream_) {
    BGNN_CHECK_ARG(A_dev && W_dev && C_dev, "null argument");
    BGNN_CHECK_ARG(heads >= 1 && N % heads == 0, "heads must divide N");
    const int CH = N / heads;
    if (!proj_bf16_supported(N, K, CH)) { bgnn_set_error("bf16 projection N=%d K=%d CH=%d is not compiled", N, K, CH); return BGNN_E_UNSUPPORTED; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return launch_proj_bf16(A_dev, W_dev, M, K, N, C_dev, att_src_dev, att_dst_dev, alpha_dev, heads, CH, sms, (cudaStream_t)stream_);
}
