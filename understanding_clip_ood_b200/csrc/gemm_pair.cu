// CTA-pair tcgen05 GEMM for the CLIP towers:  C[M,N] = epilogue(A[M,K] · W[N,K]^T + bias)  (16-bit storage, fp32 accumulate).
//
// This is the kernel the block GEMMs (QKV, out-proj, c_fc, c_proj: >96 % of the tower FLOPs) run on.  B200-first design,
// nothing of it exists in the reference (which calls F.linear, transformer.py:224-263):
//   * thread-block clusters of 2 CTAs on one TPC drive ONE 256 x BLOCK_N UMMA tile (tcgen05.mma.cta_group::2): each CTA
//     stages its own 128 rows of A and its own half of the W tile, so every operand byte is fetched into shared memory once
//     per pair (shared-memory traffic per MAC is 2/3 of the single-CTA 128 x 256 tile that saturated at ~74 % tensor-active);
//   * persistent: one cluster per SM pair, static tile walk in L2-friendly groups of M-tiles;
//   * 12 warps with fixed roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator, warps 4..11
//     epilogue;  mbarrier pipelines: smem full/empty (TMA <-> MMA; the full barrier lives in the leader CTA and is credited by
//     both CTAs' TMA loads), TMEM full/empty (MMA <-> both CTAs' epilogues; double-buffered accumulators: the epilogue of
//     tile i overlaps the MMAs of tile i+1);
//   * epilogue: tcgen05.ld -> registers -> bias / GELU / residual -> 128B-swizzled shared-memory staging -> TMA store, so
//     global writes are full 128-byte lines issued by the copy engine instead of per-thread 16-byte scatters.  The residual
//     operand is TMA-loaded into the same staging buffer ahead of time (two chunks ahead) and updated in place.
//     (Measured and rejected in round 2: one 32 x 64 TMA store per epilogue WARP with private staging and no named barrier —
//     20-25 % slower on every tower shape, profiles/r2_gemm_epilogue_ab.txt: four times as many TMA store instructions
//     compete with the operand loads for the one TMA unit of the SM.)
//   * stream-K for the ragged part of the tile grid: when the tile count does not fill whole rounds of the 74 clusters
//     (75 tiles for the out-proj / c_proj of a 128-image shard: 2 rounds for 1.01 rounds of work), the last full round plus
//     the remainder is cut into 74 EQUAL runs of K-blocks.  A cluster whose run starts inside a tile dumps that accumulator as
//     an fp32 partial into the caller's workspace and raises a flag; the cluster that holds the tile's first K-block adds the
//     partials (fixed order: deterministic) and runs the normal epilogue.  The dump comes first in a run, the fix-up last, so
//     nobody ever waits on a cluster that could be waiting itself.
// Rounding points mirror the reference's bf16/fp16 eager path: (acc + bias) is rounded to the storage type before the
// activation / residual add, which round again.
#include "common.cuh"
#include "internal.h"
#include "tmap.h"

#include <cstdlib>
#include <mutex>

// 0 (default): GELU / QuickGELU are applied to the fp32 linear output (one rounding; 3 fewer instructions per output pair in the
// epilogue that bounds the c_fc GEMM); 1: the linear output is first rounded to the storage type, the rounding point of the
// reference's 16-bit eager path (F.linear returns a 16-bit tensor) -- two roundings, further from the fp32 result
#ifndef B2C_ACT_ROUND
#define B2C_ACT_ROUND 0
#endif
// 1 (default): ONE staging buffer per epilogue group for the epilogues without a loaded residual (32 KB of staging instead of
// 64 KB; the group synchronises twice per chunk: buffer drained / buffer filled); 0: two buffers, one barrier per chunk
#ifndef B2C_STG_SINGLE
#define B2C_STG_SINGLE 1
#endif
// measurement-only builds (tools/build_variant.py): 1 = epilogue without the staging writes and the TMA store (results are
// dropped), 2 = epilogue without the TMEM load (garbage results): which part of the epilogue slows the mainloop down?
#ifndef B2C_EPI_DBG
#define B2C_EPI_DBG 0
#endif
// 1: the LN-fold epilogues stage the tile's colsum / bias vectors in shared memory once per tile (all epilogue threads load one
// element each before they wait for the accumulator) and read them with broadcast ld.shared in the chunk loop; 0: per-thread
// __ldg of the vectors inside the chunk loop (2.76 M global-load sectors per c_fc launch, profiles/r1_gemm_pair_cfc_ncu.md)
#ifndef B2C_LN_VEC_SMEM
#define B2C_LN_VEC_SMEM 1
#endif

namespace b200clip {

namespace {

constexpr int kBM = 128;          // rows of A per CTA; a pair tile has 256
constexpr int kPairM = 256;
constexpr int kBK = 64;           // 64 x 16-bit = one 128 B swizzle row
constexpr int kUmmaK = 16;
constexpr int kAccStages = 2;
constexpr int kAccStride = 256;   // TMEM columns between the two accumulator stages
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;      // two groups of 4 warps (one warp per TMEM lane quarter)
constexpr int kChunkN = 64;       // epilogue / store granularity: 128 rows x 64 columns (128 B rows)
constexpr int kChunkBytes = kBM * kChunkN * 2;
// stream-K workspace (per cluster slot): the pair tile's accumulator as fp32, laid out [CTA half][column / 4][row][4] so that
// a warp's 32 lanes (= 32 rows) touch 512 contiguous bytes per access, + one flag word per (CTA half, epilogue warp)
constexpr int kSkSlotFloat4 = 2 * (256 / 4) * kBM;
constexpr int kSkFlagsPerSlot = 2 * kEpiWarps;
// EPI values beyond the public ones: 5 = residual that aliases C (x += A W^T + b): the add is done by the L2 through a TMA
// reduce-add store, so the residual never travels to the SM (no load, no shared-memory pass)
constexpr int kEpiResidualInPlace = 5;
// 6, 7, 8 = "LN-fold" + {bias, GELU, QuickGELU}: the GEMM runs on the UN-normalised rows x and the LayerNorm is applied to
// the accumulator:  LN(x) W^T + b = rstd_m * (x W'^T)[m,n] - rstd_m * mean_m * c[n] + b'[n]   with  W' = W diag(gamma),
// c[n] = sum_k W'[n,k],  b' = b + W beta  (prepared once on the host; c, b' fp32).  Per-row (mean, rstd) come from
// row_stats_kernel.  The normalised activations never exist in memory.
constexpr int kEpiLnFold = 6;
// 9 = token-layout patch embedding: C[m,n] = round(acc) + round(pos[m % period, n]) with an fp32 table whose row 0 already
// holds class_embedding + positional_embedding[0] (the class-token rows of A are all zero), transformer.py:602-609
constexpr int kEpiPosAdd = 9;
// 10 = residual (TMA-loaded, also when it aliases C) + LayerNorm statistics of the OUTPUT rows: every epilogue thread sums
// x and x^2 of the rounded values it stores (its row, its chunks of this N tile) and writes the pair to
// stats_out[row][n_tile * 2 + group].  The next LN-fold GEMM adds the slots of a row up and derives (mean, rstd) itself, so
// the separate row-statistics pass over the residual stream (one more read of x per LayerNorm) disappears.  Deterministic:
// fixed slots, fixed summation order, no atomics.
constexpr int kEpiResidualStats = 10;
// 11 = bias + ReLU, 12 = residual (TMA-loaded) + ReLU: conv + folded BatchNorm (+ identity) + ReLU of the ModifiedResNet
// bottlenecks (deps/open_clip/src/open_clip/modified_resnet.py:42-55), the convolutions being GEMMs over NHWC rows
constexpr int kEpiRelu = 11;
constexpr int kEpiResidualRelu = 12;
constexpr bool epi_is_ln_fold(int epi) { return epi >= kEpiLnFold && epi < kEpiPosAdd; }
constexpr int epi_vec_bytes(int epi, int block_n) { return (B2C_LN_VEC_SMEM && epi_is_ln_fold(epi)) ? kAccStages * 2 * block_n * 4 : 0; }
constexpr bool epi_loads_residual(int epi) { return epi == 3 || epi == kEpiResidualStats || epi == kEpiResidualRelu; }

struct PairParams {
    const void* bias;      // storage-type bias [N]; LN-fold epilogues: fp32 b'[N]
    const float* colsum;   // LN-fold: c[N]
    const float2* rowstats; // LN-fold: (mean, rstd) per row of A ...
    const float2* stats_part; // ... or, when stats_slots > 0, `stats_slots` partial (sum x, sum x^2) pairs per row (epilogue 10)
    float2* stats_out;      // epilogue 10: [M][stats_slots] partial sums of the output rows
    int stats_slots;
    float ln_eps;
    const float* pos;      // pos-add: fp32 table [period, N]
    int pos_period;
    int M, N, K;
    int m_tiles, n_tiles;  // in units of (PAIRS*256) x BLOCK_N cluster tiles
    int group_m;
    int pf_dist;           // L2 prefetch distance of the A operand, in K-blocks (0 = off)
    int dbg;               // measurement-only switches (B200CLIP_GEMM_DBG): 1 = skip the epilogue, 2 = MMA without operand loads
    int stages;            // operand-ring depth actually used (<= the compiled depth; B200CLIP_GEMM_STAGES, measurements only)
    // stream-K: tiles [0, sk_tiles) are cut into equal runs of K-blocks over the clusters, tiles [sk_tiles, ...) stay whole
    int sk_tiles;
    float4* sk_partial;    // [clusters][kSkSlotFloat4]
    uint32_t* sk_flags;    // [clusters][kSkFlagsPerSlot], zero between launches (owners reset what they consume)
    // N-split tail (MODE 2): tiles [ns_begin, ...) -- the ragged last round -- are cut along N into ns_split pieces of
    // BLOCK_N / ns_split columns, one piece per cluster; no partial sums, no workspace (tmap_r = the W map with the piece-sized box)
    int ns_begin, ns_split;
    // implicit 3x3 convolution (MAJ bit 2): A is the NHWC activation tensor read through a 4-D tensor map, an M tile of 128 rows is
    // a block of cv_bb images x cv_by rows x cv_bx pixels, K block kb = tap (kb / cv_cblocks) x 64-channel block (kb % cv_cblocks)
    int cv_bx, cv_by, cv_bb, cv_tx, cv_ty, cv_cblocks;
    // implicit patch embedding (MAJ bit 3): A is the NCHW image batch read through a 5-D tensor map; the pixel blocks above are blocks
    // of PATCHES (cv_bx x cv_by patches of cv_bb images), K block kb = channel kb / pe_kbpc, patch rows [(kb % pe_kbpc) * pe_rows, +pe_rows)
    int pe_kbpc, pe_rows, pe_grid, pe_patch;
};

// STG_BUFS = staging buffers per epilogue group (3 with a loaded residual: landing / in-place update / store draining)
template <int BLOCK_N, int STG_BUFS, int VEC_BYTES = 0> struct PairCfg {
    static_assert(BLOCK_N % 64 == 0 && BLOCK_N >= 128 && BLOCK_N <= 256, "BLOCK_N must be 128, 192 or 256");
    static constexpr int kABytes = kBM * kBK * 2;
    static constexpr int kBBytes = (BLOCK_N / 2) * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStgBufs = STG_BUFS;
    static constexpr int kStagingBytes = 2 * kStgBufs * kChunkBytes;
    static constexpr int kBarBytes = 256;
    static constexpr int kVecBytes = VEC_BYTES;   // LN-fold: colsum | bias of the tile, one copy per accumulator stage
    static constexpr int kBudget = 227 * 1024 - 1024 - kStagingBytes - kBarBytes - kVecBytes;
    static constexpr int kStagesFit = kBudget / kStageBytes;
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kVecBytes + kBarBytes + 1024;
    static constexpr int kChunks = BLOCK_N / kChunkN;
    static_assert(kStages >= 3, "not enough shared memory for the operand ring");
    static_assert((2 * kStages + 2 * kAccStages + 2 * kStgBufs) * 8 + 8 <= kBarBytes, "barrier block too small");
};

__device__ __forceinline__ void pair_tile_coords(int t, const PairParams& p, int& mt, int& nt) {
    const int per_group = p.group_m * p.n_tiles;
    const int g = t / per_group;
    const int r = t - g * per_group;
    const int m0 = g * p.group_m;
    const int gm = min(p.group_m, p.m_tiles - m0);
    nt = r / gm;
    mt = m0 + (r - nt * gm);
}

// implicit convolution: pixel block `blk` (128 output rows) -> coordinates of its first pixel
__device__ __forceinline__ void conv_block_coords(int blk, const PairParams& p, int& x0, int& y0, int& b0) {
    const int tx = blk % p.cv_tx;
    const int r = blk / p.cv_tx;
    const int ty = r % p.cv_ty;
    x0 = tx * p.cv_bx;
    y0 = ty * p.cv_by;
    b0 = (r / p.cv_ty) * p.cv_bb;
}

// Work list of one cluster, generated identically by its producer, MMA and epilogue warps: first its run of the stream-K
// region (units = K-blocks of tiles [0, sk_tiles), run c = [c U / C, (c + 1) U / C)), then whole tiles round-robin.
struct Piece {
    int tile, kb0, kb1;
    int c0, nc;   // MODE 2 only: first 64-column chunk of the tile this piece covers, number of chunks
};
__device__ __forceinline__ int64_t sk_begin(int c, int64_t units, int clusters) { return static_cast<int64_t>(c) * units / clusters; }
// MODE 0: the plain persistent tile walk t = cluster, cluster + stride, ... (what the kernel ran before stream-K existed;
// instantiated separately so that whole-tile GEMMs carry none of the stream-K code)
template <int MODE> struct PieceIter;   // 0 = whole tiles, 1 = stream-K, 2 = whole tiles + N-split tail
template <> struct PieceIter<0> {
    int t_dp, num_tiles, num_kb, stride;
    __device__ __forceinline__ PieceIter(const PairParams& p, int num_kb_, int cluster_id, int num_clusters, int)
        : t_dp(cluster_id), num_tiles(p.m_tiles * p.n_tiles), num_kb(num_kb_), stride(num_clusters) {}
    __device__ __forceinline__ bool next(Piece& pc) {
        if (t_dp >= num_tiles) return false;
        pc.tile = t_dp;
        pc.kb0 = 0;
        pc.kb1 = num_kb;
        t_dp += stride;
        return true;
    }
};
template <> struct PieceIter<1> {
    int64_t u, end;
    int t_dp, num_tiles, num_kb, stride;
    __device__ __forceinline__ PieceIter(const PairParams& p, int num_kb_, int cluster_id, int num_clusters, int) {
        const int64_t units = static_cast<int64_t>(p.sk_tiles) * num_kb_;
        u = sk_begin(cluster_id, units, num_clusters);
        end = sk_begin(cluster_id + 1, units, num_clusters);
        t_dp = p.sk_tiles + cluster_id;
        num_tiles = p.m_tiles * p.n_tiles;
        num_kb = num_kb_;
        stride = num_clusters;
    }
    __device__ __forceinline__ bool next(Piece& pc) {
        if (u < end) {
            pc.tile = static_cast<int>(u / num_kb);
            pc.kb0 = static_cast<int>(u - static_cast<int64_t>(pc.tile) * num_kb);
            const int64_t len = min(static_cast<int64_t>(num_kb - pc.kb0), end - u);
            pc.kb1 = pc.kb0 + static_cast<int>(len);
            u += len;
            return true;
        }
        if (t_dp < num_tiles) {
            pc.tile = t_dp;
            pc.kb0 = 0;
            pc.kb1 = num_kb;
            t_dp += stride;
            return true;
        }
        return false;
    }
};

// MODE 2: whole tiles up to ns_begin (= the full rounds), then ONE piece of the N-split tail per cluster: tile ns_begin + j / s,
// chunks [(j % s) * chunks / s, ...) for cluster j < (tiles - ns_begin) * s.  Instantiated separately so that neither the whole-tile
// kernel nor the stream-K kernel carries the run-time tile width.
template <> struct PieceIter<2> {
    int t_dp, num_tiles, num_kb, stride, chunks, ns_begin, ns_split, ns_piece;
    __device__ __forceinline__ PieceIter(const PairParams& p, int num_kb_, int cluster_id, int num_clusters, int chunks_)
        : t_dp(cluster_id), num_tiles(p.ns_begin), num_kb(num_kb_), stride(num_clusters), chunks(chunks_), ns_begin(p.ns_begin),
          ns_split(p.ns_split) {
        ns_piece = cluster_id < (p.m_tiles * p.n_tiles - p.ns_begin) * p.ns_split ? cluster_id : -1;
    }
    __device__ __forceinline__ bool next(Piece& pc) {
        pc.kb0 = 0;
        pc.kb1 = num_kb;
        if (t_dp < num_tiles) {
            pc.tile = t_dp;
            pc.c0 = 0;
            pc.nc = chunks;
            t_dp += stride;
            return true;
        }
        if (ns_piece >= 0) {
            pc.tile = ns_begin + ns_piece / ns_split;
            pc.nc = chunks / ns_split;
            pc.c0 = (ns_piece % ns_split) * pc.nc;
            ns_piece = -1;
            return true;
        }
        return false;
    }
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// bounded like mbar_wait: a protocol bug becomes a trap the host reports, not a hung GPU
__device__ __forceinline__ void sk_wait_flag(const uint32_t* flag) {
    if (ld_acquire_gpu(flag) != 0u) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu(flag) == 0u) {
        if (clock64() - t0 > (1ll << 31)) {
            printf("b200clip: stream-K partial never arrived (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// The two rare stream-K paths are kept out of line (and narrow: 16 columns per step) so that they cannot raise the register
// allocation or disturb the scheduling of the epilogue's hot loop.  `tmem_row` = this warp's lane quarter of the accumulator
// stage, `slot` = this lane's row of the partial slot(s) (layout: kSkSlotFloat4 above).
__device__ __noinline__ void sk_dump_accumulator(uint32_t tmem_row, float4* slot, int grp, int chunks) {
    for (int c = grp; c < chunks; c += 2) {
#pragma unroll 1
        for (int s16 = 0; s16 < kChunkN / 16; ++s16) {
            uint32_t v[16];
            tmem_ld_32x16(tmem_row + c * kChunkN + s16 * 16, v);
            tmem_ld_wait();
            float4* dst = slot + (c * (kChunkN / 4) + s16 * 4) * kBM;
#pragma unroll
            for (int g = 0; g < 4; ++g)
                __stcg(dst + g * kBM, make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                                  __uint_as_float(v[4 * g + 3])));
        }
    }
}
// TMEM[this warp's slice] += partial slots [peer_begin, peer_end), in that order
__device__ __noinline__ void sk_add_partials(uint32_t tmem_row, const float4* slots_lane, int peer_begin, int peer_end, int grp, int chunks) {
    for (int c = grp; c < chunks; c += 2) {
#pragma unroll 1
        for (int s16 = 0; s16 < kChunkN / 16; ++s16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_row + c * kChunkN + s16 * 16;
            tmem_ld_32x16(taddr, v);
            const int off = (c * (kChunkN / 4) + s16 * 4) * kBM;
            float4 pv[4];
            {
                const float4* src = slots_lane + static_cast<int64_t>(peer_begin) * kSkSlotFloat4 + off;
#pragma unroll
                for (int g = 0; g < 4; ++g) pv[g] = __ldcg(src + g * kBM);
            }
            tmem_ld_wait();
            for (int pc2 = peer_begin;;) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    v[4 * g + 0] = __float_as_uint(__uint_as_float(v[4 * g + 0]) + pv[g].x);
                    v[4 * g + 1] = __float_as_uint(__uint_as_float(v[4 * g + 1]) + pv[g].y);
                    v[4 * g + 2] = __float_as_uint(__uint_as_float(v[4 * g + 2]) + pv[g].z);
                    v[4 * g + 3] = __float_as_uint(__uint_as_float(v[4 * g + 3]) + pv[g].w);
                }
                if (++pc2 >= peer_end) break;
                const float4* src = slots_lane + static_cast<int64_t>(pc2) * kSkSlotFloat4 + off;
#pragma unroll
                for (int g = 0; g < 4; ++g) pv[g] = __ldcg(src + g * kBM);
            }
            tmem_st_32x16(taddr, v);
        }
    }
    tmem_st_wait();
}

// PAIRS = CTA pairs per cluster (cluster size = 2 * PAIRS).  PAIRS == 2: the two pairs own vertically adjacent 256-row
// tiles of the same BLOCK_N columns; every CTA fetches one QUARTER of the W tile and TMA-multicasts it to the CTA of the
// other pair that needs the same half, so the cluster reads each W byte from L2 once instead of twice (the single-pair
// kernel is bound by the ~10 TB/s L2 -> SM read bandwidth, not by the tensor pipe: profiles/r1_gemm_pair_v1_ncu.md).
// MAJ: operand storage.  0 = both operands K-major (A [M, K], W [N, K]: the forward GEMMs).  Bit 0: A is MN-major (stored [K, M],
// M contiguous); bit 1: W is MN-major (stored [K, N]).  The backward GEMMs contract over an index that is NOT contiguous in
// memory: dgrad dX = G W reads W [N_out, K_in] as the MN-major operand (MAJ 2), wgrad dW = G^T X reads both G [tokens, N_out] and
// X [tokens, K_in] MN-major (MAJ 3) -- no transposed copies.  TMA brings MN-major tiles as [64 K rows][64 MN elements] boxes.
// Bit 2 (MAJ 4): implicit 3x3 / stride 1 / padding 1 convolution over an NHWC tensor: the A tile of K block (tap, channel block) is
// the activation tensor itself, read through a 4-D tensor map at the tap's pixel offset (out-of-image pixels arrive as zeros = the
// padding), and C is stored through the same kind of map -- the im2col matrix never exists (modified_resnet.py:17,42-47).
// Bit 3 (MAJ 8): implicit patch embedding (conv1 of the ViT, kernel = stride = patch, transformer.py:602-609): the A tile is a block
// of patches of the NCHW image batch read through a 5-D tensor map, the token rows 1 .. G*G of every image are stored through a 4-D
// map (column, patch column, patch row, image) and the positional table row of each token is added in the epilogue.
template <typename T, int BLOCK_N, int EPI, int PAIRS, int MODE, int MAJ = 0>
__global__ void __launch_bounds__(kThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r, const PairParams p) {
    constexpr bool SK = MODE == 1;                   // stream-K pieces (partial accumulators through the workspace)
    constexpr bool NS = MODE == 2;                   // N-split tail pieces (narrower UMMA on the ragged last round)
    constexpr bool kRes = epi_loads_residual(EPI);   // residual operand TMA-loaded into the staging ring
    constexpr bool kRelu = EPI == kEpiRelu || EPI == kEpiResidualRelu;
    constexpr bool kStats = EPI == kEpiResidualStats;
    using Cfg = PairCfg<BLOCK_N, kRes ? 3 : (B2C_STG_SINGLE ? 1 : 2), epi_vec_bytes(EPI, BLOCK_N)>;
    using H = Half16<T>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kStgBufs = Cfg::kStgBufs;
    constexpr bool kLn = EPI >= kEpiLnFold && EPI < kEpiPosAdd;
    constexpr bool kPos = EPI == kEpiPosAdd;
    constexpr int kAct = kLn ? EPI - kEpiLnFold : (EPI == 1 || EPI == 2 ? EPI : 0);  // 0 none, 1 GELU, 2 QuickGELU
    constexpr int kClusterCtas = 2 * PAIRS;
    constexpr int kClusterM = PAIRS * kPairM;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024 B alignment.  The dynamic-smem base offset is the same in every CTA of the cluster, so the
    // carve-up below puts every object at the same CTA-relative address in all of them (required by the pair MMA, the
    // multicast TMA and the multicast commits).
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint8_t* staging = smem + kStages * Cfg::kStageBytes;
    float* vec_smem = reinterpret_cast<float*>(staging + Cfg::kStagingBytes);   // [acc stage][colsum | bias][BLOCK_N] (LN-fold only)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes + Cfg::kVecBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + kAccStages;
    uint64_t* res_bar = tmem_empty_bar + kAccStages;  // [group][buffer]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + 2 * kStgBufs);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const uint32_t pair = cta_rank >> 1;           // which pair of the cluster
    const uint32_t half = cta_rank & 1;            // which CTA of the pair (0 = leader: issues the MMAs, owns the barriers)
    const uint32_t leader_rank = cta_rank & ~1u;
    const bool is_leader = half == 0;
    const int cluster_id = blockIdx.x / kClusterCtas;
    const int num_clusters = gridDim.x / kClusterCtas;
    const int num_tiles = p.m_tiles * p.n_tiles;
    const int num_kb = (p.K + kBK - 1) / kBK;
    const int ring = p.stages > 0 && p.stages < kStages ? p.stages : kStages;

    // every CTA of the cluster must be resident before the pair-wide TMEM allocation / remote barrier traffic
    cluster_sync_all();
    pdl_launch_dependents();  // the next kernel's CTAs may take over this SM as soon as this CTA retires

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_c);
        if constexpr (kRes) tma_prefetch_desc(&tmap_r);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 2);       // leader: arrive.expect_tx, peer: remote arrive (+ the TMA bytes of both CTAs' buffers)
            mbar_init(&empty_bar[i], PAIRS);  // one multicast tcgen05.commit from every pair leader of the cluster
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(&tmem_full_bar[i], 1);                // multicast tcgen05.commit of this pair's leader
            mbar_init(&tmem_empty_bar[i], 2 * kEpiWarps);   // every epilogue warp of both CTAs arrives at the leader
        }
        for (int i = 0; i < 2 * kStgBufs; ++i) mbar_init(&res_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_ptr_smem, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // everything above overlapped the previous kernel's tail; its outputs (our A operand / residual) are needed from here on
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer (every CTA) =====================
        if (lane == 0 && !(p.dbg & 2)) {
            int stage = 0;
            uint32_t phase = 0;
            const int row_in_cluster = static_cast<int>(pair) * kPairM + static_cast<int>(half) * kBM;
            // L2 prefetch cursor (whole-tile schedules only): runs p.pf_dist K-blocks ahead of the shared-memory ring (across tile
            // boundaries), so the ring only has to cover L2 latency, not the DRAM latency of the streamed activations
            const int pf_dist = (MODE != 0 || MAJ != 0) ? 0 : p.pf_dist;
            int pf_t = cluster_id, pf_kb = 0, pf_row = 0;
            auto prefetch_next = [&]() {
                if (pf_t >= num_tiles) return;
                if (pf_kb == 0) {
                    int pmt, pnt;
                    pair_tile_coords(pf_t, p, pmt, pnt);
                    pf_row = pmt * kClusterM + row_in_cluster;
                }
                tma_prefetch_l2_2d(&tmap_a, pf_kb * kBK, pf_row);
                if (++pf_kb == num_kb) {
                    pf_kb = 0;
                    pf_t += num_clusters;
                }
            };
            for (int i = 0; i < pf_dist; ++i) prefetch_next();
            PieceIter<MODE> pieces(p, num_kb, cluster_id, num_clusters, Cfg::kChunks);
            Piece pc;
            while (pieces.next(pc)) {
                int mt, nt;
                pair_tile_coords(pc.tile, p, mt, nt);
                const int row_a = mt * kClusterM + row_in_cluster;
                int cv_x0 = 0, cv_y0 = 0, cv_b0 = 0;
                if constexpr ((MAJ & 12) != 0) conv_block_coords(row_a / kBM, p, cv_x0, cv_y0, cv_b0);
                // a piece of the N-split tail covers pc.nc of the tile's 64-column chunks: this CTA holds half of those W rows
                const bool narrow = NS && pc.nc != Cfg::kChunks;
                const int piece_w_rows = NS ? pc.nc * (kChunkN / 2) : BLOCK_N / 2;
                const int row_w = nt * BLOCK_N + (NS ? pc.c0 * kChunkN : 0) + static_cast<int>(half) * piece_w_rows;
                const uint32_t stage_tx = NS ? 2u * static_cast<uint32_t>(Cfg::kABytes + piece_w_rows * kBK * 2) : 2u * Cfg::kStageBytes;
                for (int kb = pc.kb0; kb < pc.kb1; ++kb) {
                    if (pf_dist > 0) prefetch_next();
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                    else mbar_arrive_remote(&full_bar[stage], leader_rank);
                    if constexpr ((MAJ & 8) != 0) {
                        // implicit patch embedding: this K block is channel kb / kbpc, pe_rows pixel rows of every patch of my block
                        // (one box per pixel row: 128 patches x P pixels, P * 2 bytes = one swizzle row of the narrow layout)
                        const int ch = kb / p.pe_kbpc, dy0 = (kb - ch * p.pe_kbpc) * p.pe_rows;
                        for (int rr = 0; rr < p.pe_rows; ++rr)
                            tma_load_5d_pair(&tmap_a, &full_bar[stage], smem_a + stage * Cfg::kABytes + rr * (kBM * p.pe_patch * 2), 0, dy0 + rr, cv_x0,
                                             cv_y0, cv_b0 * 3 + ch, kCacheHintEvictNormal);
                    } else if constexpr ((MAJ & 4) != 0) {
                        // implicit 3x3 convolution: this K block is tap kb / cblocks, channels [64 (kb % cblocks), +64) of my pixel block
                        const int tap = kb / p.cv_cblocks, cb = kb - tap * p.cv_cblocks;
                        tma_load_4d_pair(&tmap_a, &full_bar[stage], smem_a + stage * Cfg::kABytes, cb * kBK, cv_x0 + tap % 3 - 1, cv_y0 + tap / 3 - 1,
                                         cv_b0, kCacheHintEvictNormal);
                    } else if constexpr ((MAJ & 1) != 0) {
                        // MN-major A: my 128 rows of the tile = two [64 k][64 m] boxes
#pragma unroll
                        for (int blk = 0; blk < kBM / 64; ++blk)
                            tma_load_2d_pair(&tmap_a, &full_bar[stage], smem_a + stage * Cfg::kABytes + blk * 8192, row_a + blk * 64, kb * kBK,
                                             kCacheHintEvictNormal);
                    } else {
                        tma_load_2d_pair(&tmap_a, &full_bar[stage], smem_a + stage * Cfg::kABytes, kb * kBK, row_a, kCacheHintEvictNormal);
                    }
                    if constexpr ((MAJ & 2) != 0) {
                        static_assert(MAJ == 0 || (PAIRS == 1 && MODE != 2 && (BLOCK_N / 2) % 64 == 0), "MN-major W: single pairs, whole 64-column blocks");
#pragma unroll
                        for (int blk = 0; blk < BLOCK_N / 128; ++blk)
                            tma_load_2d_pair(&tmap_w, &full_bar[stage], smem_b + stage * Cfg::kBBytes + blk * 8192, row_w + blk * 64, kb * kBK,
                                             kCacheHintEvictLast);
                    } else if constexpr (PAIRS == 1) {
                        if (narrow)   // tmap_r = the W map with the piece-sized box (MODE 2 never loads a residual)
                            tma_load_2d_pair(&tmap_r, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBK, row_w, kCacheHintEvictLast);
                        else
                            tma_load_2d_pair(&tmap_w, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBK, row_w, kCacheHintEvictLast);
                    } else {
                        // my quarter of the W tile -> the CTAs holding W half `half` in both pairs
                        constexpr int kQRows = BLOCK_N / 4;
                        const uint16_t mask = static_cast<uint16_t>((1u << half) | (1u << (half + 2)));
                        tma_load_2d_pair_mcast(&tmap_w, &full_bar[stage], smem_b + stage * Cfg::kBBytes + pair * (kQRows * kBK * 2),
                                               kb * kBK, row_w + static_cast<int>(pair) * kQRows, mask, kCacheHintEvictLast);
                    }
                    if (++stage == ring) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (pair leaders only) =====================
        if (is_leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(H::kUmmaFormat, kPairM, BLOCK_N) | ((MAJ & 1) ? (1u << 15) : 0u) | ((MAJ & 2) ? (1u << 16) : 0u);
            constexpr uint32_t kStepA = (MAJ & 1) ? (kUmmaK * 128) >> 4 : (kUmmaK * 2) >> 4;   // descriptor advance per UMMA K step
            constexpr uint32_t kStepB = (MAJ & 2) ? (kUmmaK * 128) >> 4 : (kUmmaK * 2) >> 4;
            constexpr uint16_t kAllCtas = static_cast<uint16_t>((1u << kClusterCtas) - 1);
            const uint16_t pair_mask = static_cast<uint16_t>(0x3u << leader_rank);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            PieceIter<MODE> pieces(p, num_kb, cluster_id, num_clusters, Cfg::kChunks);
            Piece pc;
            for (; pieces.next(pc); ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kAccStride;
                // N-split tail pieces run a narrower UMMA (N = 64 x the piece's chunks) on the same A rows
                const uint32_t idesc_pc = NS ? make_idesc_f16(H::kUmmaFormat, kPairM, static_cast<uint32_t>(pc.nc * kChunkN)) : idesc;
                for (int kb = pc.kb0; kb < pc.kb1; ++kb) {
                    if (!(p.dbg & 2)) mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t addr_a = smem_u32(smem_a + stage * Cfg::kABytes), addr_b = smem_u32(smem_b + stage * Cfg::kBBytes);
                    const uint64_t desc_a = (MAJ & 1) ? make_sw128_mnmajor_desc_lbo(addr_a, 8192) : make_sw128_kmajor_desc(addr_a);
                    const uint64_t desc_b = (MAJ & 2) ? make_sw128_mnmajor_desc_lbo(addr_b, 8192) : make_sw128_kmajor_desc(addr_b);
                    if constexpr ((MAJ & 8) != 0) {
                        // patch embedding: the A stage is 64 / P sub-tiles of [128 patches][P pixels] in the narrow swizzle
                        const uint32_t row_bytes = static_cast<uint32_t>(p.pe_patch) * 2u;
#pragma unroll
                        for (int k = 0; k < kBK / kUmmaK; ++k) {
                            const uint32_t e = static_cast<uint32_t>(k * kUmmaK);
                            const uint32_t sub = e / static_cast<uint32_t>(p.pe_patch), off = (e - sub * p.pe_patch) * 2u;
                            umma_f16_pair(tmem_d, make_narrow_kmajor_desc(addr_a + sub * (kBM * row_bytes) + off, row_bytes), desc_b + kStepB * k,
                                          idesc_pc, ((kb - pc.kb0) | k) != 0);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < kBK / kUmmaK; ++k)
                            umma_f16_pair(tmem_d, desc_a + kStepA * k, desc_b + kStepB * k, idesc_pc, ((kb - pc.kb0) | k) != 0);
                    }
                    // frees the slot in EVERY CTA of the cluster (each of them writes into some of the buffers just read)
                    if (!(p.dbg & 2)) umma_commit_pair(&empty_bar[stage], kAllCtas);
                    if (kb == pc.kb1 - 1) umma_commit_pair(&tmem_full_bar[acc], pair_mask);
                    if (++stage == ring) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            // drain: the peer's last remote arrivals must have landed before this CTA may exit
            if (it > 0) {
                const int last = it - 1;
                mbar_wait(&tmem_empty_bar[last & 1], (last >> 1) & 1);
                if (it > 1) {
                    const int prev = it - 2;
                    mbar_wait(&tmem_empty_bar[prev & 1], (prev >> 1) & 1);
                }
            }
        }
        __syncwarp();
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue (every CTA) =====================
        const int e = warp - kEpiWarp0;
        const int q = warp & 3;   // TMEM lane quarter this warp may access
        const int grp = e >> 2;   // epilogue group: handles column chunks grp, grp+2, ...
        const bool grp_leader = (e & 3) == 0 && lane == 0;
        const int bar_id = 1 + grp;
        const int r = q * 32 + lane;  // row inside this CTA's 128-row half
        uint8_t* stg_ptr = staging + grp * kStgBufs * kChunkBytes;
        const uint32_t stg_base = smem_u32(stg_ptr);
        uint64_t* my_res_bar = res_bar + grp * kStgBufs;
        const uint32_t row_off = static_cast<uint32_t>(r) * 128;
        const uint32_t rx = static_cast<uint32_t>(r & 7);
        const int row_in_cluster = static_cast<int>(pair) * kPairM + static_cast<int>(half) * kBM;
        const T* bias = static_cast<const T*>(p.bias);
        const float* bias_f32 = static_cast<const float*>(p.bias);
        uint32_t bufc = 0;  // chunks processed by this group so far (buffer = bufc % kStgBufs)
#if B2C_EPI_DBG == 1
        uint32_t dbg_sink = 0;
#endif
        // stream-K bookkeeping of this warp: its part of every cluster slot, its flag in every slot
        const int64_t sk_units = static_cast<int64_t>(p.sk_tiles) * num_kb;
        const int sk_lane_off = static_cast<int>(half) * (BLOCK_N / 4) * kBM + r;       // + col4 * kBM, in float4
        const int sk_flag_off = static_cast<int>(half) * kEpiWarps + e;

        // (tile, chunk) -> the next chunk this group processes (whole-tile schedules; the residual prefetch runs ahead on it)
        auto advance = [&](int& tile, int& chunk) {
            chunk += 2;
            if (chunk >= Cfg::kChunks) {
                chunk = grp;
                tile += num_clusters;
            }
        };
        auto issue_residual = [&](int tile, int chunk, uint32_t use) {  // group leader only
            int mt2, nt2;
            pair_tile_coords(tile, p, mt2, nt2);
            const uint32_t b = use % kStgBufs;
            mbar_arrive_expect_tx(&my_res_bar[b], kChunkBytes);
            tma_load_2d(&tmap_r, &my_res_bar[b], stg_ptr + b * kChunkBytes, nt2 * BLOCK_N + chunk * kChunkN,
                        mt2 * kClusterM + row_in_cluster, kCacheHintEvictFirst);
        };
        // residual prefetch, two chunks ahead: the first two chunks of this group
        if constexpr (kRes) {
            if (grp_leader && !(p.dbg & 1)) {
                int pt = cluster_id, pcn = grp;
                if (pt < num_tiles) issue_residual(pt, pcn, 0);
                advance(pt, pcn);
                if (pt < num_tiles) issue_residual(pt, pcn, 1);
            }
            __syncwarp();
        }

        int it = 0;
        PieceIter<MODE> pieces(p, num_kb, cluster_id, num_clusters, Cfg::kChunks);
        Piece pc;
        for (; pieces.next(pc); ++it) {
            const int t = pc.tile;
            int mt, nt;
            pair_tile_coords(t, p, mt, nt);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int row0 = mt * kClusterM + row_in_cluster;
            int cv_x0 = 0, cv_y0 = 0, cv_b0 = 0;
            if constexpr ((MAJ & 12) != 0) conv_block_coords(row0 / kBM, p, cv_x0, cv_y0, cv_b0);
            const int pc_nc = NS ? pc.nc : Cfg::kChunks;   // 64-column chunks of this piece (N-split tail: fewer than the tile's)
            const int pc_c0 = NS ? pc.c0 : 0;
            const bool dump = SK && pc.kb0 != 0;                          // run starts inside the tile: accumulator -> fp32 partial
            const bool fixup = SK && pc.kb0 == 0 && pc.kb1 < num_kb;     // tile's first K-block, but not its last: add the others' partials

            if (SK && dump) {
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                tc_fence_after();
                if (!(p.dbg & 1))
                    sk_dump_accumulator(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride,
                                        p.sk_partial + static_cast<int64_t>(cluster_id) * kSkSlotFloat4 + sk_lane_off, grp, Cfg::kChunks);
                tc_fence_before();
                __threadfence();      // every lane's partial rows are visible device-wide before the flag goes up
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_remote(&tmem_empty_bar[acc], leader_rank);
                    st_release_gpu(p.sk_flags + cluster_id * kSkFlagsPerSlot + sk_flag_off, 1u);
                }
                __syncwarp();
                continue;
            }

            // clusters (cluster_id, peer_end) hold the rest of this tile's K range, one partial each
            int peer_end = cluster_id + 1;
            if (fixup) {
                const int64_t tile_end = static_cast<int64_t>(t + 1) * num_kb;
                while (peer_end < num_clusters && sk_begin(peer_end, sk_units, num_clusters) < tile_end) ++peer_end;
                if (lane == 0 && !(p.dbg & 1))
                    for (int pc2 = cluster_id + 1; pc2 < peer_end; ++pc2) sk_wait_flag(p.sk_flags + pc2 * kSkFlagsPerSlot + sk_flag_off);
                __syncwarp();
            }

            uint32_t vec_addr = 0;
            if constexpr (kLn && Cfg::kVecBytes > 0) {
                // the tile's colsum | bias -> shared memory, one element per epilogue thread, while the MMAs of this tile still run.
                // Buffer `acc` was last read two tiles ago: every epilogue warp has passed the barrier of the tile in between.
                float* vs = vec_smem + acc * 2 * BLOCK_N;
                const int et = static_cast<int>(threadIdx.x) - kEpiWarp0 * 32;
                if (et < BLOCK_N) {
                    const int ecol = nt * BLOCK_N + et;
                    const bool eok = ecol < p.N;
                    vs[et] = eok ? __ldg(p.colsum + ecol) : 0.f;
                    vs[BLOCK_N + et] = eok ? __ldg(bias_f32 + ecol) : 0.f;
                }
                named_bar_sync(3, kEpiWarps * 32);
                vec_addr = smem_u32(vs);
            }
            const float* pos_row = nullptr;
            if constexpr (kPos && (MAJ & 8) != 0) {
                // row r of my block is patch (cv_y0 + yl, cv_x0 + xl) of image cv_b0 + bl: token 1 + py * G + px (patches beyond the
                // grid are clipped by the store: any valid table row will do for them)
                const int xl = r % p.cv_bx, yl = (r / p.cv_bx) % p.cv_by;
                const int px = cv_x0 + xl, py = cv_y0 + yl;
                const int token = (px < p.pe_grid && py < p.pe_grid) ? 1 + py * p.pe_grid + px : 0;
                pos_row = p.pos + static_cast<int64_t>(token) * p.N;
            } else if constexpr (kPos) pos_row = p.pos + static_cast<int64_t>((row0 + r) % p.pos_period) * p.N;
            float ln_rstd = 0.f, ln_nmr = 0.f;
            if constexpr (kLn) {
                if (row0 + r < p.M) {
                    float2 st;
                    if (p.stats_slots > 0) {
                        // partial (sum, sum of squares) pairs written by the epilogue-10 GEMM that produced these rows
                        const float2* sp = p.stats_part + static_cast<int64_t>(row0 + r) * p.stats_slots;
                        float sx = 0.f, sq = 0.f;
                        for (int i = 0; i < p.stats_slots; ++i) {
                            const float2 v2 = __ldg(sp + i);
                            sx += v2.x;
                            sq += v2.y;
                        }
                        const float inv_k = 1.0f / static_cast<float>(p.K);
                        const float mean = sx * inv_k;
                        st.x = mean;
                        st.y = rsqrtf(fmaxf(fmaf(-mean, mean, sq * inv_k), 0.f) + p.ln_eps);
                    } else {
                        st = __ldg(p.rowstats + row0 + r);
                    }
                    ln_rstd = st.y;
                    ln_nmr = -st.x * st.y;
                }
            }
            float st_s = 0.f, st_q = 0.f;   // epilogue 10: this thread's share of sum x / sum x^2 of its output row
            if constexpr (kRes) {
                // pull the residual tiles this group will need two tiles from now into L2
                if (grp_leader && p.pf_dist > 0) {
                    const int ft = t + 2 * num_clusters;
                    if (ft < num_tiles) {
                        int fm, fn;
                        pair_tile_coords(ft, p, fm, fn);
                        for (int c = grp; c < Cfg::kChunks; c += 2)
                            tma_prefetch_l2_2d(&tmap_r, fn * BLOCK_N + c * kChunkN, fm * kClusterM + row_in_cluster);
                    }
                }
                __syncwarp();
            }
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            if (p.dbg & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[acc], leader_rank);
                continue;
            }
            if (fixup) {
                // Stream-K fix-up as a separate pass over this warp's slice of the accumulator: TMEM += the partial sums of
                // the tile's remaining K range, in cluster order (deterministic).  The ordinary epilogue below then runs
                // unchanged (its loop is the hot path of every GEMM: nothing of the fix-up lives in it).
                sk_add_partials(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride, p.sk_partial + sk_lane_off,
                                cluster_id + 1, peer_end, grp, Cfg::kChunks);
                // this warp has consumed its part of every peer's partial: the flags are zero again for the next launch
                __syncwarp();
                if (lane == 0)
                    for (int pc2 = cluster_id + 1; pc2 < peer_end; ++pc2) p.sk_flags[pc2 * kSkFlagsPerSlot + sk_flag_off] = 0u;
                __syncwarp();
            }

            if (NS && grp >= pc_nc) {
                // a one-chunk piece of the N-split tail leaves the second epilogue group without work: free the accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[acc], leader_rank);
                __syncwarp();
                continue;
            }
#pragma unroll 1
            for (int c = grp; c < pc_nc; c += 2) {
                const bool last_of_tile = c + 2 >= pc_nc;
                const uint32_t buf = bufc % kStgBufs;
                const uint32_t stg = stg_base + buf * kChunkBytes;
                // Staging buffer `buf` is free here: the group leader drains its outstanding TMA store before it joins the
                // barrier that ends each chunk (below), so the store issued kStgBufs chunks ago finished reading long ago.
                if constexpr (kRes) mbar_wait(&my_res_bar[buf], (bufc / kStgBufs) & 1);
                const int col0 = nt * BLOCK_N + (pc_c0 + c) * kChunkN;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccStride + c * kChunkN + hf * 32;
#if B2C_EPI_DBG == 2
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(ln_rstd * static_cast<float>(i + lane));
#else
                    tmem_ld_32x32(taddr, v);
#endif
                    uint4 bvec[4];
                    if constexpr (!kLn) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int col = col0 + hf * 32 + g * 8;
                            bvec[g] = make_uint4(0, 0, 0, 0);
                            if (bias != nullptr && col < p.N) bvec[g] = ldg128(bias + col);
                        }
                    }
                    tmem_ld_wait();
                    if constexpr (!kRes && kStgBufs == 1) {
                        if (hf == 0) {
                            // single staging buffer: the previous chunk's store must have finished reading it
                            if (grp_leader) tma_store_wait_read<0>();
                            __syncwarp();
                            named_bar_sync(bar_id, 128);
                        }
                    }
                    if (last_of_tile && hf == 1) {
                        // last TMEM read of this tile: hand the accumulator stage back to the MMA issuer early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(&tmem_empty_bar[acc], leader_rank);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t saddr = stg + row_off + (((static_cast<uint32_t>(hf * 4 + g)) ^ rx) << 4);
                        const uint32_t bw[4] = {bvec[g].x, bvec[g].y, bvec[g].z, bvec[g].w};
                        uint32_t rw[4] = {0, 0, 0, 0};
                        if constexpr (kRes) {
                            const uint4 rv = lds128(saddr);
                            rw[0] = rv.x; rw[1] = rv.y; rw[2] = rv.z; rw[3] = rv.w;
                        }
                        uint32_t ow[4];
                        float cf[8], bf[8];
                        if constexpr (kPos) {
                            const int col = col0 + hf * 32 + g * 8;
                            const bool ok = col < p.N;
#pragma unroll
                            for (int q4 = 0; q4 < 2; ++q4) {
                                const float4 b4 = ok ? __ldg(reinterpret_cast<const float4*>(pos_row + col) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
                                bf[q4 * 4 + 0] = b4.x; bf[q4 * 4 + 1] = b4.y; bf[q4 * 4 + 2] = b4.z; bf[q4 * 4 + 3] = b4.w;
                            }
                        }
                        if constexpr (kLn && Cfg::kVecBytes > 0) {
                            // broadcast reads of the staged vectors (every lane the same address)
                            const uint32_t va = vec_addr + static_cast<uint32_t>(((pc_c0 + c) * kChunkN + hf * 32 + g * 8) * 4);
#pragma unroll
                            for (int q4 = 0; q4 < 2; ++q4) {
                                const uint4 c4 = lds128(va + q4 * 16);
                                const uint4 b4 = lds128(va + BLOCK_N * 4 + q4 * 16);
                                cf[q4 * 4 + 0] = __uint_as_float(c4.x); cf[q4 * 4 + 1] = __uint_as_float(c4.y);
                                cf[q4 * 4 + 2] = __uint_as_float(c4.z); cf[q4 * 4 + 3] = __uint_as_float(c4.w);
                                bf[q4 * 4 + 0] = __uint_as_float(b4.x); bf[q4 * 4 + 1] = __uint_as_float(b4.y);
                                bf[q4 * 4 + 2] = __uint_as_float(b4.z); bf[q4 * 4 + 3] = __uint_as_float(b4.w);
                            }
                        } else if constexpr (kLn) {
                            const int col = col0 + hf * 32 + g * 8;
                            const bool ok = col < p.N;
#pragma unroll
                            for (int q4 = 0; q4 < 2; ++q4) {
                                const float4 c4 = ok ? __ldg(reinterpret_cast<const float4*>(p.colsum + col) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
                                const float4 b4 = ok ? __ldg(reinterpret_cast<const float4*>(bias_f32 + col) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
                                cf[q4 * 4 + 0] = c4.x; cf[q4 * 4 + 1] = c4.y; cf[q4 * 4 + 2] = c4.z; cf[q4 * 4 + 3] = c4.w;
                                bf[q4 * 4 + 0] = b4.x; bf[q4 * 4 + 1] = b4.y; bf[q4 * 4 + 2] = b4.z; bf[q4 * 4 + 3] = b4.w;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x0, x1;
                            if constexpr (kLn) {
                                // rstd * acc + (b' - rstd * mean * c): two packed FMAs per pair
                                const uint64_t t = fma_f2(pack_f2(ln_nmr, ln_nmr), pack_f2(cf[2 * j], cf[2 * j + 1]), pack_f2(bf[2 * j], bf[2 * j + 1]));
                                unpack_f2(fma_f2(pack_f2(__uint_as_float(v[g * 8 + 2 * j]), __uint_as_float(v[g * 8 + 2 * j + 1])),
                                                 pack_f2(ln_rstd, ln_rstd), t), x0, x1);
                            } else if constexpr (kPos) {
                                // conv output and table entry are each rounded to the storage type before the add
                                const float2 a2 = H::unpack(H::pack(__uint_as_float(v[g * 8 + 2 * j]), __uint_as_float(v[g * 8 + 2 * j + 1])));
                                const float2 p2 = H::unpack(H::pack(bf[2 * j], bf[2 * j + 1]));
                                x0 = a2.x + p2.x;
                                x1 = a2.y + p2.y;
                            } else {
                                const float2 b2 = H::unpack(bw[j]);
                                unpack_f2(add_f2(pack_f2(__uint_as_float(v[g * 8 + 2 * j]), __uint_as_float(v[g * 8 + 2 * j + 1])),
                                                 pack_f2(b2.x, b2.y)), x0, x1);
                            }
                            if constexpr ((kAct != 0 && B2C_ACT_ROUND) || kRes || EPI == kEpiResidualInPlace) {
                                const float2 xr = H::unpack(H::pack(x0, x1));  // linear output rounded to the storage type
                                x0 = xr.x;
                                x1 = xr.y;
                            }
                            if constexpr (kAct == 1) {
                                gelu_pair_fast(x0, x1);
                            } else if constexpr (kAct == 2) {
                                x0 = quick_gelu(x0);
                                x1 = quick_gelu(x1);
                            } else if constexpr (kRes) {
                                const float2 r2 = H::unpack(rw[j]);
                                x0 += r2.x;
                                x1 += r2.y;
                            }
                            if constexpr (kRelu) {
                                x0 = fmaxf(x0, 0.f);
                                x1 = fmaxf(x1, 0.f);
                            }
                            ow[j] = H::pack(x0, x1);
                            if constexpr (kStats) {
                                const float2 rr = H::unpack(ow[j]);   // the values the next GEMM will read
                                st_s += rr.x + rr.y;
                                st_q = fmaf(rr.x, rr.x, fmaf(rr.y, rr.y, st_q));
                            }
                        }
#if B2C_EPI_DBG == 1
                        dbg_sink ^= ow[0] ^ ow[1] ^ ow[2] ^ ow[3];
#else
                        sts128(saddr, make_uint4(ow[0], ow[1], ow[2], ow[3]));
#endif
                    }
                }
#if B2C_EPI_DBG == 1
                if (dbg_sink == 0x7fc12345u && p.dbg == 77) p.sk_flags[threadIdx.x] = dbg_sink;   // keeps the arithmetic alive
                ++bufc;
                continue;
#endif
                fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
                if constexpr (!kRes && kStgBufs > 1) {
                    // every store but (at most) the previous chunk's has long completed; waiting for that one too BEFORE the
                    // barrier tells the whole group that the other staging buffer is free for the next chunk
                    if (grp_leader) tma_store_wait_read<0>();
                    __syncwarp();
                }
                named_bar_sync(bar_id, 128);
                if (grp_leader) {
                    if constexpr (EPI == kEpiResidualInPlace) tma_reduce_add_2d(&tmap_c, stg_ptr + buf * kChunkBytes, col0, row0);
                    else if constexpr ((MAJ & 12) != 0) tma_store_4d(&tmap_c, stg_ptr + buf * kChunkBytes, col0, cv_x0, cv_y0, cv_b0);
                    else tma_store_2d(&tmap_c, stg_ptr + buf * kChunkBytes, col0, row0);
                    tma_store_commit();
                    if constexpr (kRes) {
                        // prefetch the residual of the chunk two steps ahead into the buffer whose store was committed one
                        // step ago (everything but the store just committed must have released its buffer)
                        int pt = t, pcn = c;
                        advance(pt, pcn);
                        advance(pt, pcn);
                        if (pt < num_tiles) {
                            tma_store_wait_read<1>();
                            issue_residual(pt, pcn, bufc + 2);
                        }
                    }
                }
                __syncwarp();
                ++bufc;
            }
            if constexpr (kStats) {
                if (row0 + r < p.M)
                    p.stats_out[static_cast<int64_t>(row0 + r) * p.stats_slots + nt * 2 + grp] = make_float2(st_s, st_q);
            }
        }
        if (grp_leader) tma_store_wait_all<0>();  // global writes complete before the CTA retires
        __syncwarp();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// clusters the stream-K schedule is planned for = one per SM pair of the device (the persistent grid of a busy GEMM)
int sk_clusters_planned() { return num_sms() / 2; }

template <typename T, int BLOCK_N, int EPI, int PAIRS, int MODE, int MAJ = 0>
int launch_pair_sk(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const CUtensorMap& tr, const PairParams& p,
                   cudaStream_t stream) {
    using Cfg = PairCfg<BLOCK_N, epi_loads_residual(EPI) ? 3 : (B2C_STG_SINGLE ? 1 : 2), epi_vec_bytes(EPI, BLOCK_N)>;
    auto kern = gemm_pair_kernel<T, BLOCK_N, EPI, PAIRS, MODE, MAJ>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    static int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * PAIRS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (attr_err != cudaSuccess) return;
        // how many clusters of this shape the device can hold at once (clusters of 4 do not tile all 148 SMs)
        cfg.gridDim = dim3(num_sms() / (2 * PAIRS) * (2 * PAIRS));
        int n = 0;
        attr_err = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        max_clusters = n;
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "gemm_pair: kernel attribute / cluster occupancy query");
    B2C_CHECK_ARG(max_clusters > 0, "gemm_pair: the device cannot co-schedule a cluster of %d CTAs", 2 * PAIRS);
    const int tiles = p.m_tiles * p.n_tiles;
    int clusters = tiles < max_clusters ? tiles : max_clusters;
    if constexpr (MODE != 0) {
        // the stream-K split / the N-split tail were planned for `sk_clusters` clusters (stream-K: all co-resident, flag waits)
        B2C_CHECK_ARG(sk_clusters_planned() <= max_clusters, "gemm_pair: stream-K needs %d co-resident clusters, device holds %d",
                      sk_clusters_planned(), max_clusters);
        clusters = sk_clusters_planned();
    }
    cfg.gridDim = dim3(2 * PAIRS * clusters);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tw, tc, tr, p);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm_pair_kernel)");
    B2C_LAUNCH_CHECK("gemm_pair_kernel");
    return 0;
}

// whole-tile GEMMs run the kernel instantiated without any stream-K code; the stream-K instantiation exists for single pairs and
// the epilogues without a TMA-loaded residual
template <typename T, int BLOCK_N, int EPI, int PAIRS>
int launch_pair(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const CUtensorMap& tr, const PairParams& p,
                cudaStream_t stream) {
    if constexpr (PAIRS == 1 && !epi_loads_residual(EPI)) {
        if (p.sk_tiles > 0) return launch_pair_sk<T, BLOCK_N, EPI, PAIRS, 1>(ta, tw, tc, tr, p, stream);
        if (p.ns_split > 0) return launch_pair_sk<T, BLOCK_N, EPI, PAIRS, 2>(ta, tw, tc, tr, p, stream);
    }
    B2C_CHECK_ARG(p.sk_tiles == 0 && p.ns_split == 0, "gemm_pair: stream-K / N-split is not available for this kernel variant");
    return launch_pair_sk<T, BLOCK_N, EPI, PAIRS, 0>(ta, tw, tc, tr, p, stream);
}

template <typename T, int BLOCK_N, int PAIRS>
int launch_pair_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const CUtensorMap& tr,
                    const PairParams& p, cudaStream_t s) {
    switch (epi) {
        case 0: return launch_pair<T, BLOCK_N, 0, PAIRS>(ta, tw, tc, tr, p, s);
        case 1: return launch_pair<T, BLOCK_N, 1, PAIRS>(ta, tw, tc, tr, p, s);
        case 2: return launch_pair<T, BLOCK_N, 2, PAIRS>(ta, tw, tc, tr, p, s);
        case 3: return launch_pair<T, BLOCK_N, 3, PAIRS>(ta, tw, tc, tr, p, s);
        case kEpiResidualInPlace: return launch_pair<T, BLOCK_N, kEpiResidualInPlace, PAIRS>(ta, tw, tc, tr, p, s);
        case kEpiLnFold + 0: return launch_pair<T, BLOCK_N, kEpiLnFold + 0, PAIRS>(ta, tw, tc, tr, p, s);
        case kEpiLnFold + 1: return launch_pair<T, BLOCK_N, kEpiLnFold + 1, PAIRS>(ta, tw, tc, tr, p, s);
        case kEpiLnFold + 2: return launch_pair<T, BLOCK_N, kEpiLnFold + 2, PAIRS>(ta, tw, tc, tr, p, s);
        case kEpiPosAdd: return launch_pair<T, BLOCK_N, kEpiPosAdd, PAIRS>(ta, tw, tc, tr, p, s);
        case kEpiResidualStats:
            if constexpr (PAIRS == 1) return launch_pair<T, BLOCK_N, kEpiResidualStats, 1>(ta, tw, tc, tr, p, s);
            break;
        case kEpiRelu:
            if constexpr (PAIRS == 1) return launch_pair<T, BLOCK_N, kEpiRelu, 1>(ta, tw, tc, tr, p, s);
            break;
        case kEpiResidualRelu:
            if constexpr (PAIRS == 1) return launch_pair<T, BLOCK_N, kEpiResidualRelu, 1>(ta, tw, tc, tr, p, s);
            break;
    }
    set_last_error("gemm_pair: unsupported epilogue %d", epi);
    return -1;
}

template <typename T, int PAIRS>
int launch_pair_bn(int bn, int epi, const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const CUtensorMap& tr,
                   const PairParams& p, cudaStream_t s) {
    switch (bn) {
        case 256: return launch_pair_epi<T, 256, PAIRS>(epi, ta, tw, tc, tr, p, s);
        case 192: return launch_pair_epi<T, 192, PAIRS>(epi, ta, tw, tc, tr, p, s);
        case 128: return launch_pair_epi<T, 128, PAIRS>(epi, ta, tw, tc, tr, p, s);
    }
    set_last_error("gemm_pair: BLOCK_N must be 128, 192 or 256 (got %d)", bn);
    return -1;
}

// B200CLIP_GEMM_PAIRS=1|2 overrides the cluster shape (A/B measurements); default: single pairs
int default_gemm_pairs(int M) {
    static const int forced = [] {
        const char* e = getenv("B200CLIP_GEMM_PAIRS");
        return e != nullptr ? atoi(e) : 0;
    }();
    if (forced == 1 || forced == 2) return forced;
    (void)M;
    return 1;
}

// B200CLIP_NO_REDUCE_STORE=1: in-place residual GEMMs load the residual instead of using the TMA reduce-add store
bool no_reduce_store() {
    static const bool v = [] {
        const char* e = getenv("B200CLIP_NO_REDUCE_STORE");
        return e != nullptr && e[0] == '1';
    }();
    return v;
}

int gemm_debug_switches() {
    static const int v = [] {
        const char* e = getenv("B200CLIP_GEMM_DBG");
        return e != nullptr ? atoi(e) : 0;
    }();
    return v;
}

// B200CLIP_GEMM_STAGES=<n> shortens the operand ring (ring-depth sensitivity measurements: 3 -> 4 stages is worth 3 %, 4 -> 5
// and 5 -> 6 nothing, profiles/r2_gemm_epilogue_dissection.txt)
int gemm_ring_override() {
    static const int v = [] {
        const char* e = getenv("B200CLIP_GEMM_STAGES");
        return e != nullptr ? atoi(e) : 0;
    }();
    return v;
}

// B200CLIP_PF_DIST=<K-blocks> overrides the L2 prefetch distance of the A operand (0 disables it)
int l2_prefetch_distance() {
    static const int v = [] {
        const char* e = getenv("B200CLIP_PF_DIST");
        const int d = e != nullptr ? atoi(e) : 0;
        return d < 0 ? 0 : (d > 64 ? 64 : d);
    }();
    return v;
}

}  // namespace

// N-split of the ragged last round: `rem` tiles of `bn` columns left over after `full` whole rounds are cut along N into `s` pieces
// each (piece width bn / s, a multiple of the 64-column epilogue chunk), one piece per cluster, when all pieces fit into one round.
// Measured (profiles/r2_gemm_epilogue_dissection.txt (6)): worth 2-3 % on the short grids of a 128-image shard (c_fc 33.3 -> 32.4 us,
// ViT-B/32 forward 1.61 -> 1.57 ms) and a LOSS on long grids / long K (the few clusters of a ragged last round have the whole L2
// bandwidth to themselves, so that round is much shorter than a full one, while narrow pieces re-read A once per piece): only for
// grids of at most 6 whole rounds and K < 32 K-blocks.  Returns s (0 = leave the tail as whole tiles).
// B200CLIP_NSPLIT=0 disables it, =1 applies it whenever the pieces fit.
static int plan_n_split(int full, int rem, int bn, int clusters, int num_kb = 0) {
    static const int mode = [] {
        const char* e = getenv("B200CLIP_NSPLIT");
        return e != nullptr ? atoi(e) : -1;
    }();
    if (mode == 0 || rem <= 0) return 0;
    const double eff = static_cast<double>(full * clusters + rem) / (static_cast<double>(full + 1) * clusters);
    if (mode != 1 && (eff >= 0.93 || full > 6 || num_kb >= 32)) return 0;
    const int chunks = bn / kChunkN;
    for (int s = chunks; s >= 2; --s)
        if (chunks % s == 0 && rem * s <= clusters) return s;
    return 0;
}
// measured cost per MAC relative to the 256-wide tile (profiles/r1_gemm_variants_vs_cublas.txt; 64: the N-split tail pieces):
// narrower tiles read more operand bytes per MAC through the L2 -> SM path that bounds the mainloop
static double width_cost(int w) { return w >= 256 ? 1.0 : w >= 192 ? 1.18 : w >= 128 ? 1.45 : 2.2; }

// Tile-shape choice: the persistent grid runs ceil(tiles / clusters) rounds; pick the N tile that minimises
// rounds x tile cost (a 256-wide tile is the most efficient per MAC, narrower ones waste less of the last round).
// stream_k: the ragged part will be evened out by stream-K (long K); n_split: the last round may be cut along N.
int pick_pair_block_n(int M, int N, int pairs, bool stream_k = false, bool n_split = false, int num_kb = 0) {
    const int clusters = pairs == 2 ? 33 : num_sms() / 2;
    const long mt = (M + pairs * kPairM - 1) / (pairs * kPairM);
    double best_cost = 1e30;
    int best = 256;
    const int cands[3] = {256, 192, 128};
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        const long nt = (N + bn - 1) / bn;
        const long tiles = mt * nt;
        const int full = static_cast<int>(tiles / clusters);
        const int rem = static_cast<int>(tiles % clusters);
        // whole rounds, or -- with stream-K evening out the ragged part -- the exact share of work per cluster
        double cost = static_cast<double>(full + (rem > 0 ? 1 : 0)) * bn * width_cost(bn);
        if (stream_k && tiles > clusters) {
            cost = static_cast<double>(tiles) / clusters * bn * width_cost(bn);
        } else if (n_split && rem > 0) {
            const int s = plan_n_split(full, rem, bn, clusters, num_kb);
            if (s > 0) cost = static_cast<double>(full) * bn * width_cost(bn) + (bn / s) * width_cost(bn / s);
        }
        if (cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

// Partial-sum slots per row that an epilogue-10 GEMM of this shape writes (2 per N tile); the consumer passes it back.
int gemm_pair_stats_slots(int M, int N) { return 2 * ((N + pick_pair_block_n(M, N, 1) - 1) / pick_pair_block_n(M, N, 1)); }

// Stream-K workspace: one fp32 accumulator slot + flag words per cluster.  Fixed size (independent of the problem).
int64_t gemm_pair_sk_workspace_bytes() {
    const int64_t clusters = sk_clusters_planned();
    return clusters * kSkSlotFloat4 * 16 + ((clusters * kSkFlagsPerSlot * 4 + 255) / 256) * 256;
}
static uint32_t* sk_flags_of(void* ws) {
    return reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + static_cast<int64_t>(sk_clusters_planned()) * kSkSlotFloat4 * 16);
}
// flags must be zero before the first stream-K GEMM on a workspace (afterwards the kernels leave them at zero)
int gemm_pair_sk_workspace_reset(void* ws, cudaStream_t stream) {
    B2C_CHECK_ARG(ws != nullptr, "gemm: null stream-K workspace");
    B2C_CUDA(cudaMemsetAsync(sk_flags_of(ws), 0, static_cast<size_t>(sk_clusters_planned()) * kSkFlagsPerSlot * 4, stream));
    return 0;
}

// How many tiles (from the front of the tile order) go through stream-K: none when whole rounds fit (or nearly: >= 94 % of
// the last round busy), otherwise the remainder plus — when there is one — one full round, so that every cluster gets
// between one and two tiles' worth of K-blocks and no tile is cut into more than two or three pieces.
// B200CLIP_STREAMK=0 disables it, =1 forces it whenever there is a remainder (A/B measurements).
static int plan_stream_k(int tiles, int num_kb, int clusters) {
    static const int mode = [] {
        const char* e = getenv("B200CLIP_STREAMK");
        return e != nullptr ? atoi(e) : -1;
    }();
    if (mode == 0 || clusters <= 1) return 0;
    // the dump + fix-up of a split tile cost about as much as 8 K-blocks of tensor work: worth it only for long K
    // (c_proj, K = 4 W: 48 K-blocks), not for the K = W GEMMs (12) -- profiles/r2_gemm_stream_k.txt
    if (mode != 1 && num_kb < 32) return 0;
    const int rem = tiles % clusters;
    if (rem == 0) return 0;
    const int full = tiles / clusters;
    const double eff = static_cast<double>(tiles) / (static_cast<double>(full + 1) * clusters);
    if (mode != 1 && eff >= 0.94) return 0;
    const int sk = rem + (full >= 1 ? clusters : 0);
    // every cluster must get a run of at least 4 K-blocks (pieces shorter than the operand ring are all fill and drain)
    if (static_cast<int64_t>(sk) * num_kb < static_cast<int64_t>(clusters) * 4) return 0;
    return sk;
}

// pairs: 1 = clusters of 2 CTAs, 2 = clusters of 4 with W multicast, 0 = choose
int gemm_pair(bool is_bf16, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* residual,
              int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue, int force_block_n, int pairs,
              cudaStream_t stream, const float* ln_colsum, const float* ln_rowstats, const float* pos_table, int pos_period,
              float* stats_out, const float* stats_part, int stats_slots, float ln_eps, void* sk_workspace) {
    B2C_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    B2C_CHECK_ARG((epilogue >= 0 && epilogue <= 3) || epilogue == 5 || epilogue == 6, "gemm_pair: unsupported epilogue %d", epilogue);
    const bool relu = epilogue == 5 || epilogue == 6;          // B200CLIP_EPI_RELU / B200CLIP_EPI_RESIDUAL_RELU
    if (epilogue == 5) epilogue = 0;
    if (epilogue == 6) epilogue = 3;
    const bool ln = ln_colsum != nullptr || ln_rowstats != nullptr || stats_part != nullptr;
    if (ln) {
        B2C_CHECK_ARG(ln_colsum != nullptr && (ln_rowstats != nullptr) != (stats_part != nullptr) && bias != nullptr && epilogue <= 2,
                      "gemm_ln: needs colsum, rowstats OR partial sums, an fp32 bias and a bias/GELU/QuickGELU epilogue");
        B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(ln_colsum) | reinterpret_cast<uintptr_t>(bias)) % 16 == 0 &&
                          (reinterpret_cast<uintptr_t>(ln_rowstats) | reinterpret_cast<uintptr_t>(stats_part)) % 8 == 0,
                      "gemm_ln: colsum / bias must be 16-byte aligned, rowstats 8-byte aligned");
        B2C_CHECK_ARG(stats_part == nullptr || (stats_slots > 0 && stats_slots <= 64), "gemm_ln: bad partial-sum slot count %d", stats_slots);
    }
    if (stats_out != nullptr)
        B2C_CHECK_ARG(epilogue == 3 && !ln && pos_table == nullptr && reinterpret_cast<uintptr_t>(stats_out) % 8 == 0,
                      "gemm_stats: row statistics are produced by the residual epilogue only");
    B2C_CHECK_ARG(N % 8 == 0 || (bias == nullptr && epilogue == 0), "gemm: N=%d must be a multiple of 8 (unless bias-free)", N);
    B2C_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0,
                  "gemm: K, lda, ldw, ldc must be multiples of 8 (16 B rows) K=%d lda=%lld ldw=%lld ldc=%lld", K,
                  (long long)lda, (long long)ldw, (long long)ldc);
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) % 16 == 0,
                  "gemm: A, W, C must be 16-byte aligned");
    B2C_CHECK_ARG(bias == nullptr || reinterpret_cast<uintptr_t>(bias) % 16 == 0, "gemm: bias must be 16-byte aligned");
    if (epilogue == 3) {
        B2C_CHECK_ARG(residual != nullptr && ldr % 8 == 0 && reinterpret_cast<uintptr_t>(residual) % 16 == 0,
                      "gemm: residual epilogue needs an aligned residual pointer");
    }
    B2C_CHECK_ARG(pairs >= 0 && pairs <= 2, "gemm_pair: pairs must be 0, 1 or 2");
    if (pairs == 0) pairs = default_gemm_pairs(M);
    if (stats_out != nullptr) pairs = 1;
    // stream-K needs the caller's workspace, single pairs, and an epilogue without the TMA-loaded residual (whose prefetch runs
    // ahead on the whole-tile order): the in-place residual form (TMA reduce-add store) qualifies, a separate residual does not
    const bool res_in_place = epilogue == 3 && !relu && residual == C && ldr == ldc && !no_reduce_store();
    if (relu) pairs = 1;
    const bool sk_variant = pairs == 1 && stats_out == nullptr && (epilogue != 3 || res_in_place);   // MODE 1 / 2 instantiations exist
    const bool sk_ok = sk_workspace != nullptr && sk_variant;
    // stream-K pays off for long K only (plan_stream_k); the K = W GEMMs get their ragged last round cut along N instead
    const bool long_k = (K + kBK - 1) / kBK >= 32;
    const int bn = force_block_n > 0 ? force_block_n : pick_pair_block_n(M, N, pairs, sk_ok && long_k, sk_variant, (K + kBK - 1) / kBK);

    CUtensorMap ta, tw, tc, tr;
    if (make_tmap_2d(&ta, is_bf16, A, M, K, lda, kBM, kBK) != 0) return -1;
    if (make_tmap_2d(&tw, is_bf16, W, N, K, ldw, pairs == 2 ? bn / 4 : bn / 2, kBK) != 0) return -1;
    if (make_tmap_2d(&tc, is_bf16, C, M, N, ldc, kBM, kChunkN) != 0) return -1;
    if (stats_out != nullptr) {
        if (make_tmap_2d(&tr, is_bf16, residual, M, N, ldr, kBM, kChunkN) != 0) return -1;
        epilogue = kEpiResidualStats;
    } else if (res_in_place) {
        epilogue = kEpiResidualInPlace;
        tr = tc;
    } else if (epilogue == 3) {
        if (make_tmap_2d(&tr, is_bf16, residual, M, N, ldr, kBM, kChunkN) != 0) return -1;
    } else {
        tr = tc;
    }

    if (relu) {
        B2C_CHECK_ARG(!ln && pos_table == nullptr && stats_out == nullptr, "gemm_pair: the ReLU epilogues take a plain bias (+ residual)");
        epilogue = epilogue == 3 ? kEpiResidualRelu : kEpiRelu;
    }
    if (ln) epilogue += kEpiLnFold;
    if (pos_table != nullptr) {
        B2C_CHECK_ARG(!ln && epilogue == 0 && bias == nullptr && pos_period > 0 && reinterpret_cast<uintptr_t>(pos_table) % 16 == 0,
                      "gemm_pos: needs a bias-free plain epilogue, an aligned fp32 table and period > 0");
        epilogue = kEpiPosAdd;
    }
    PairParams p;
    p.bias = bias;
    p.colsum = ln_colsum;
    p.rowstats = reinterpret_cast<const float2*>(ln_rowstats);
    p.stats_part = reinterpret_cast<const float2*>(stats_part);
    p.stats_out = reinterpret_cast<float2*>(stats_out);
    p.stats_slots = stats_out != nullptr ? 2 * ((N + bn - 1) / bn) : (stats_part != nullptr ? stats_slots : 0);
    p.ln_eps = ln_eps;
    p.pos = pos_table;
    p.pos_period = pos_period;
    p.M = M;
    p.N = N;
    p.K = K;
    p.m_tiles = (M + pairs * kPairM - 1) / (pairs * kPairM);
    p.n_tiles = (N + bn - 1) / bn;
    p.group_m = pairs == 2 ? 4 : 8;
    p.pf_dist = l2_prefetch_distance();
    p.dbg = gemm_debug_switches();
    p.stages = gemm_ring_override();
    p.sk_tiles = 0;
    p.sk_partial = nullptr;
    p.sk_flags = nullptr;
    p.ns_begin = 0;
    p.ns_split = 0;
    p.cv_bx = p.cv_by = p.cv_bb = p.cv_tx = p.cv_ty = p.cv_cblocks = 0;
    p.pe_kbpc = p.pe_rows = p.pe_grid = p.pe_patch = 0;
    if (sk_ok && epilogue != 3 && epilogue != kEpiResidualStats) {
        B2C_CHECK_ARG(reinterpret_cast<uintptr_t>(sk_workspace) % 16 == 0, "gemm: stream-K workspace must be 16-byte aligned");
        p.sk_tiles = plan_stream_k(p.m_tiles * p.n_tiles, (K + kBK - 1) / kBK, sk_clusters_planned());
        p.sk_partial = static_cast<float4*>(sk_workspace);
        p.sk_flags = sk_flags_of(sk_workspace);
    }
    if (sk_variant && p.sk_tiles == 0 && epilogue != 3 && epilogue != kEpiResidualStats && epilogue != kEpiResidualRelu) {
        // ragged last round: cut its tiles along N, one piece per cluster (no partial sums, no workspace)
        const int tiles = p.m_tiles * p.n_tiles, clusters = sk_clusters_planned();
        const int split = plan_n_split(tiles / clusters, tiles % clusters, bn, clusters, (K + kBK - 1) / kBK);
        if (split > 0) {
            p.ns_split = split;
            p.ns_begin = tiles / clusters * clusters;
            // MODE 2 never loads a residual: its fourth tensor map carries W with the piece-sized box
            if (make_tmap_2d(&tr, is_bf16, W, N, K, ldw, bn / split / 2, kBK) != 0) return -1;
        }
    }
    if (pairs == 2)
        return is_bf16 ? launch_pair_bn<__nv_bfloat16, 2>(bn, epilogue, ta, tw, tc, tr, p, stream)
                       : launch_pair_bn<__half, 2>(bn, epilogue, ta, tw, tc, tr, p, stream);
    return is_bf16 ? launch_pair_bn<__nv_bfloat16, 1>(bn, epilogue, ta, tw, tc, tr, p, stream)
                   : launch_pair_bn<__half, 1>(bn, epilogue, ta, tw, tc, tr, p, stream);
}

// Backward-GEMM form: C[M, N] = op(A) op(W)^T with plain stores (no bias / activation), operands optionally MN-major (see the MAJ
// template parameter): a_mn: A is stored [K, M] (row pitch lda), w_mn: W is stored [K, N] (row pitch ldw).  dgrad = (false, true),
// wgrad = (true, true).  Stream-K through `sk_workspace` as in gemm_pair (the wgrad contraction runs over all token rows: a few
// dozen output tiles with hundreds of K-blocks).
// B200CLIP_MN_WIDE=1: prefer the 256-wide tile for stream-K grids with fewer tiles than clusters (A/B measurements)
static bool mn_wide_tiles() {
    static const bool v = [] {
        const char* e = getenv("B200CLIP_MN_WIDE");
        return e != nullptr && e[0] == '1';
    }();
    return v;
}

template <typename T, int BLOCK_N, int MAJ>
static int launch_pair_mn(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const PairParams& p, cudaStream_t s) {
    if (p.sk_tiles > 0) return launch_pair_sk<T, BLOCK_N, 0, 1, 1, MAJ>(ta, tw, tc, tc, p, s);
    return launch_pair_sk<T, BLOCK_N, 0, 1, 0, MAJ>(ta, tw, tc, tc, p, s);
}
template <typename T>
static int launch_pair_mn_bn(int bn, int maj, const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const PairParams& p,
                             cudaStream_t s) {
    if (bn == 256 && maj == 2) return launch_pair_mn<T, 256, 2>(ta, tw, tc, p, s);
    if (bn == 256 && maj == 3) return launch_pair_mn<T, 256, 3>(ta, tw, tc, p, s);
    if (bn == 128 && maj == 2) return launch_pair_mn<T, 128, 2>(ta, tw, tc, p, s);
    if (bn == 128 && maj == 3) return launch_pair_mn<T, 128, 3>(ta, tw, tc, p, s);
    set_last_error("gemm_pair_mn: unsupported tile / operand layout (BLOCK_N %d, layout %d)", bn, maj);
    return -1;
}

int gemm_pair_mn(bool is_bf16, const void* A, int64_t lda, bool a_mn, const void* W, int64_t ldw, bool w_mn, void* C, int64_t ldc, int M, int N,
                 int K, cudaStream_t stream, void* sk_workspace) {
    B2C_CHECK_ARG(M > 0 && N > 0 && K > 0 && A && W && C, "gemm_mn: bad problem M=%d N=%d K=%d", M, N, K);
    B2C_CHECK_ARG(w_mn, "gemm_mn: W must be MN-major (the all-K-major form is gemm_pair)");
    B2C_CHECK_ARG(N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && (a_mn ? M % 8 == 0 : K % 8 == 0),
                  "gemm_mn: N, the row pitches and the contiguous extent of A must be multiples of 8 (16 B rows)");
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) % 16 == 0,
                  "gemm_mn: A, W, C must be 16-byte aligned");
    const int num_kb = (K + kBK - 1) / kBK;
    const int clusters = sk_clusters_planned();
    const bool sk_ok = sk_workspace != nullptr;
    // N tile: 256 or 128 (an MN-major W tile is made of whole 64-column blocks per CTA), same cost model as gemm_pair
    int bn = 256;
    {
        double best = 1e30;
        const int cands[2] = {256, 128};
        const long mt = (M + kPairM - 1) / kPairM;
        for (int i = 0; i < 2; ++i) {
            const long tiles = mt * ((N + cands[i] - 1) / cands[i]);
            double rounds = static_cast<double>((tiles + clusters - 1) / clusters);
            // long K with the workspace: stream-K gives every cluster an equal share also when there are FEWER tiles than clusters
            // (wgrad: 27-54 tiles of 100-154 K-blocks), so the wide tile's lower operand traffic per MAC decides
            if (sk_ok && num_kb >= 32 && (tiles > clusters || mn_wide_tiles())) rounds = static_cast<double>(tiles) / clusters;
            const double cost = rounds * cands[i] * width_cost(cands[i]);
            if (cost < best) {
                best = cost;
                bn = cands[i];
            }
        }
    }
    CUtensorMap ta, tw, tc;
    if (a_mn) {
        if (make_tmap_2d(&ta, is_bf16, A, K, M, lda, 64, 64) != 0) return -1;     // [K rows][M contiguous], boxes of 64 k x 64 m
    } else {
        if (make_tmap_2d(&ta, is_bf16, A, M, K, lda, kBM, kBK) != 0) return -1;
    }
    if (make_tmap_2d(&tw, is_bf16, W, K, N, ldw, 64, 64) != 0) return -1;          // [K rows][N contiguous]
    if (make_tmap_2d(&tc, is_bf16, C, M, N, ldc, kBM, kChunkN) != 0) return -1;
    PairParams p;
    p.bias = nullptr;
    p.colsum = nullptr;
    p.rowstats = nullptr;
    p.stats_part = nullptr;
    p.stats_out = nullptr;
    p.stats_slots = 0;
    p.ln_eps = 0.f;
    p.pos = nullptr;
    p.pos_period = 0;
    p.M = M;
    p.N = N;
    p.K = K;
    p.m_tiles = (M + kPairM - 1) / kPairM;
    p.n_tiles = (N + bn - 1) / bn;
    p.group_m = 8;
    p.pf_dist = 0;
    p.dbg = gemm_debug_switches();
    p.stages = gemm_ring_override();
    p.sk_tiles = 0;
    p.sk_partial = nullptr;
    p.sk_flags = nullptr;
    p.ns_begin = 0;
    p.ns_split = 0;
    p.cv_bx = p.cv_by = p.cv_bb = p.cv_tx = p.cv_ty = p.cv_cblocks = 0;
    p.pe_kbpc = p.pe_rows = p.pe_grid = p.pe_patch = 0;
    if (sk_ok) {
        B2C_CHECK_ARG(reinterpret_cast<uintptr_t>(sk_workspace) % 16 == 0, "gemm_mn: stream-K workspace must be 16-byte aligned");
        p.sk_tiles = plan_stream_k(p.m_tiles * p.n_tiles, num_kb, clusters);
        p.sk_partial = static_cast<float4*>(sk_workspace);
        p.sk_flags = sk_flags_of(sk_workspace);
    }
    const int maj = (a_mn ? 1 : 0) | 2;
    return is_bf16 ? launch_pair_mn_bn<__nv_bfloat16>(bn, maj, ta, tw, tc, p, stream) : launch_pair_mn_bn<__half>(bn, maj, ta, tw, tc, p, stream);
}

// Implicit 3x3 / stride 1 / padding 1 convolution + bias + ReLU over NHWC activations (the conv2 of a ModifiedResNet bottleneck with
// its folded BatchNorm, modified_resnet.py:17-18,45):  out[b, y, x, n] = relu(sum_{tap, c} in[b, y + dy, x + dx, c] w[n, tap * C + c] + bias[n]).
// C must be a multiple of 64 (one K block = 64 channels of one tap); W and H must be divisible by the pixel-block edges chosen here
// (powers of two, bx * by * bb = 128) -- returns 1 ("not applicable") otherwise so that the caller falls back to im2col + GEMM.
int gemm_pair_conv3x3(bool is_bf16, const void* in, const void* Wt, const void* bias, void* out, int batch, int H, int W, int C, int N,
                      cudaStream_t stream) {
    B2C_CHECK_ARG(in && Wt && bias && out && batch > 0 && H > 0 && W > 0 && C > 0 && N > 0, "conv3x3: bad arguments");
    if (C % kBK != 0 || N % 8 != 0) return 1;
    int bx = 1, by = 1;
    while (bx < 16 && W % (bx * 2) == 0) bx *= 2;
    while (bx * by < kBM && by < 16 && H % (by * 2) == 0) by *= 2;
    if (bx * by > kBM || bx < 2 || by < 2) return 1;
    const int bb = kBM / (bx * by);
    if (bb > 256) return 1;
    const int bn = N > 128 ? 256 : 128;
    CUtensorMap ta, tw, tc;
    if (make_tmap_nhwc(&ta, is_bf16, in, batch, H, W, C, bb, by, bx) != 0) return -1;
    if (make_tmap_2d(&tw, is_bf16, Wt, N, 9 * static_cast<uint64_t>(C), 9 * static_cast<uint64_t>(C), bn / 2, kBK) != 0) return -1;
    if (make_tmap_nhwc(&tc, is_bf16, out, batch, H, W, N, bb, by, bx) != 0) return -1;
    PairParams p;
    p.bias = bias;
    p.colsum = nullptr;
    p.rowstats = nullptr;
    p.stats_part = nullptr;
    p.stats_out = nullptr;
    p.stats_slots = 0;
    p.ln_eps = 0.f;
    p.pos = nullptr;
    p.pos_period = 0;
    p.cv_bx = bx;
    p.cv_by = by;
    p.cv_bb = bb;
    p.cv_tx = W / bx;
    p.cv_ty = H / by;
    p.cv_cblocks = C / kBK;
    p.pe_kbpc = p.pe_rows = p.pe_grid = p.pe_patch = 0;
    const int64_t blocks = static_cast<int64_t>(p.cv_tx) * p.cv_ty * ((batch + bb - 1) / bb);
    B2C_CHECK_ARG(blocks * kBM < (1ll << 31), "conv3x3: too many output rows");
    p.M = static_cast<int>(blocks * kBM);          // padded row count (pixel blocks beyond the batch are clipped by the tensor maps)
    p.N = N;
    p.K = 9 * C;
    p.m_tiles = static_cast<int>((blocks + 1) / 2);
    p.n_tiles = (N + bn - 1) / bn;
    p.group_m = 8;
    p.pf_dist = 0;
    p.dbg = gemm_debug_switches();
    p.stages = gemm_ring_override();
    p.sk_tiles = 0;
    p.sk_partial = nullptr;
    p.sk_flags = nullptr;
    p.ns_begin = 0;
    p.ns_split = 0;
    if (is_bf16)
        return bn == 256 ? launch_pair_sk<__nv_bfloat16, 256, kEpiRelu, 1, 0, 4>(ta, tw, tc, tc, p, stream)
                         : launch_pair_sk<__nv_bfloat16, 128, kEpiRelu, 1, 0, 4>(ta, tw, tc, tc, p, stream);
    return bn == 256 ? launch_pair_sk<__half, 256, kEpiRelu, 1, 0, 4>(ta, tw, tc, tc, p, stream)
                     : launch_pair_sk<__half, 128, kEpiRelu, 1, 0, 4>(ta, tw, tc, tc, p, stream);
}

// Implicit patch embedding of the ViT (conv1, kernel = stride = patch, no bias, transformer.py:602-609) for 16-bit NCHW batches:
//   x[b, 1 + py * G + px, :] = round(patch(b, py, px) . conv1_w^T) + round(pos_cls[1 + py * G + px, :]),   x[b, 0, :] = round(pos_cls[0, :])
// (pos_cls row 0 already holds class_embedding + positional_embedding[0]).  The patches are read from the image batch itself through
// a 5-D tensor map -- no im2col matrix.  Patch sizes with 64 % P == 0 and 16-byte pixel rows (16, 32); returns 1 otherwise.
namespace {
template <typename T>
__global__ void __launch_bounds__(256) class_rows_kernel(T* __restrict__ x, const float* __restrict__ pos0, int batch, int64_t image_pitch, int width) {
    const int64_t total = static_cast<int64_t>(batch) * width;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int64_t b = i / width;
        const int j = static_cast<int>(i - b * width);
        const float v = pos0[j];
        x[b * image_pitch + j] = Half16<T>::from_f(v);
    }
}
}  // namespace

int gemm_pair_patch_embed(bool is_bf16, const void* image, const void* conv1_w, const float* pos_cls, void* x, int batch, int image_size,
                          int patch, int width, cudaStream_t stream) {
    B2C_CHECK_ARG(image && conv1_w && pos_cls && x && batch > 0 && image_size > 0 && patch > 0 && width > 0, "patch_embed: bad arguments");
    if (kBK % patch != 0 || (patch * 2) % 16 != 0 || image_size % patch != 0 || width % 8 != 0) return 1;
    const int G = image_size / patch, L = G * G + 1;
    int bx = 1;
    while (bx < G) bx *= 2;
    if (bx > 16) return 1;
    int by = 1;
    while (by < G && bx * by * 2 <= kBM) by *= 2;
    const int bb = kBM / (bx * by);
    const int rows = kBK / patch;
    const int K = 3 * patch * patch;
    const int bn = width > 128 ? 256 : 128;
    CUtensorMap ta, tw, tc;
    if (make_tmap_patches(&ta, is_bf16, image, batch, image_size, patch, rows, bb, by, bx) != 0) return -1;
    if (make_tmap_2d(&tw, is_bf16, conv1_w, width, K, K, bn / 2, kBK) != 0) return -1;
    // token rows 1 .. L-1 of x [batch, L, width] as (column, patch column, patch row, image)
    {
        CUtensorMap* m = &tc;
        // a 4-D map with the same box order as the NHWC maps: dims {width, G, G, batch}, strides {width, G * width, L * width}
        // (make_tmap_nhwc assumes dense packing, so this one is spelled out through the generic helper below)
        if (make_tmap_nhwc_strided(m, is_bf16, static_cast<char*>(x) + static_cast<int64_t>(width) * 2, batch, G, G, width,
                                   static_cast<uint64_t>(width), static_cast<uint64_t>(G) * width, static_cast<uint64_t>(L) * width, bb, by, bx) != 0)
            return -1;
    }
    PairParams p;
    p.bias = nullptr;
    p.colsum = nullptr;
    p.rowstats = nullptr;
    p.stats_part = nullptr;
    p.stats_out = nullptr;
    p.stats_slots = 0;
    p.ln_eps = 0.f;
    p.pos = pos_cls;
    p.pos_period = L;
    p.cv_bx = bx;
    p.cv_by = by;
    p.cv_bb = bb;
    p.cv_tx = (G + bx - 1) / bx;
    p.cv_ty = (G + by - 1) / by;
    p.cv_cblocks = 0;
    p.pe_kbpc = patch * patch / kBK;
    p.pe_rows = rows;
    p.pe_grid = G;
    p.pe_patch = patch;
    const int64_t blocks = static_cast<int64_t>(p.cv_tx) * p.cv_ty * ((batch + bb - 1) / bb);
    B2C_CHECK_ARG(blocks * kBM < (1ll << 31), "patch_embed: too many rows");
    p.M = static_cast<int>(blocks * kBM);
    p.N = width;
    p.K = K;
    p.m_tiles = static_cast<int>((blocks + 1) / 2);
    p.n_tiles = (width + bn - 1) / bn;
    p.group_m = 8;
    p.pf_dist = 0;
    p.dbg = gemm_debug_switches();
    p.stages = gemm_ring_override();
    p.sk_tiles = 0;
    p.sk_partial = nullptr;
    p.sk_flags = nullptr;
    p.ns_begin = 0;
    p.ns_split = 0;
    int rc;
    if (is_bf16)
        rc = bn == 256 ? launch_pair_sk<__nv_bfloat16, 256, kEpiPosAdd, 1, 0, 8>(ta, tw, tc, tc, p, stream)
                       : launch_pair_sk<__nv_bfloat16, 128, kEpiPosAdd, 1, 0, 8>(ta, tw, tc, tc, p, stream);
    else
        rc = bn == 256 ? launch_pair_sk<__half, 256, kEpiPosAdd, 1, 0, 8>(ta, tw, tc, tc, p, stream)
                       : launch_pair_sk<__half, 128, kEpiPosAdd, 1, 0, 8>(ta, tw, tc, tc, p, stream);
    if (rc != 0) return rc;
    const int64_t total = static_cast<int64_t>(batch) * width;
    const int grid = static_cast<int>((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    if (is_bf16)
        class_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(x), pos_cls, batch, static_cast<int64_t>(L) * width, width);
    else
        class_rows_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<__half*>(x), pos_cls, batch, static_cast<int64_t>(L) * width, width);
    B2C_LAUNCH_CHECK("class_rows_kernel");
    return 0;
}

}  // namespace b200clip
