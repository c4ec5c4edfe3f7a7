// tcgen05 attention for the longer sequences of the CLIP towers (64 < L <= 257: ViT-B/16 L=197, ViT-L/14 L=257, the
// untruncated text context L=77), head_dim 64.  Replaces the SDPA core of nn.MultiheadAttention as used by
// ResidualAttentionBlock.attention (deps/open_clip/src/open_clip/transformer.py:224,238-251), causal variant = the text
// tower's strict upper-triangular mask (transformer.py:751-757).
//
// Persistent, one CTA per SM.  A work item is one (batch, head): its K and V rows are TMA-loaded ONCE into a two-stage
// shared-memory ring (the next item's K/V arrive while this one is computed) and serve all of its 128-row query tiles.
// The whole key range of a tile fits one accumulator (<= 256 TMEM columns), so there is no online-softmax rescaling.
// Two query tiles are in flight, one per 256-column half of tensor memory, each owned by its own softmax warpgroup:
//
//   warp 8 (one thread)   TMA: K / V boxes per item, Q box [128 x 64] per tile, straight out of the packed qkv matrix
//   warp 9 (one thread)   tcgen05.mma  S[128 x n] = Q K^T   (SS; accumulator in TMEM half t&1)
//                         tcgen05.mma  O[128 x 64] = P V    (A = P read from TMEM, B = V as an MN-major smem operand: V is
//                                                            never transposed)
//   warps 0-3 / 4-7       thread = query row of an even / odd tile: pass 1 row max over tcgen05.ld chunks, pass 2
//                         p = 2^(s*c - m*c), row sum, P packed to 16 bit and written back with tcgen05.st over S columns
//                         this thread has already consumed; then O * (1/sum) -> swizzled staging tile -> TMA store.
// While one warpgroup runs the exponentials of tile t, the tensor pipe produces S of tile t+1 and the other warpgroup
// drains O of tile t-1, so the MUFU pipe (16 ex2 / clk / SM: the binding unit of this kernel) stays busy.
//
// Inside a TMEM half: S in columns [0, n), P (16-bit pairs) in [0, n/2), O in [128, 192).
// A tail tile with <= 64 valid rows is issued as M = 64 (rows 16w..16w+15 live in lanes 0-15 of quadrant w); warps whose
// rows are all beyond the sequence skip the softmax work.  L = 257 (CLS + 16 x 16 patches) keeps n = 256 on the tensor
// pipe and handles key 256 on the CUDA cores (one 64-term dot product and one 64-term axpy per row).
#include "common.cuh"
#include "internal.h"
#include "tmap.h"

#include <mutex>

namespace b200clip {

namespace {

constexpr int kHd = 64;
constexpr int kQT = 128;               // query rows per tile
constexpr int kTileQBytes = kQT * 128;
constexpr int kThreads = 384;   // 2 softmax warpgroups + 1 warpgroup holding the TMA and MMA threads
constexpr int kTmemCols = 512;
constexpr int kMaxSmem = 227 * 1024;

struct AttnTcParams {
    int L, heads, q_tiles, items, tail_rows, n_cols, kv_stage_bytes;
    int nq, nkv;   // ring depths: query-tile slots (2..4), K/V stages (2..6), as many as shared memory holds
    long long* dbg; // timeline capture of CTA 0 (tools/attn_timeline.py); nullptr in normal operation
};

// event e of tile t of CTA 0 -> dbg[t * 8 + e] (clock64)
#define B2C_STAMP(e, t) do { if (p.dbg != nullptr && blockIdx.x == 0 && (t) < 64) p.dbg[(t) * 8 + (e)] = clock64(); } while (0)

// tcgen05.mma with the A operand in tensor memory (lane = row, one 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void attn_tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// tcgen05.wait::ld + a register dependency on the loaded values: nothing may be scheduled on `v` before the wait.
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i += 8)
        asm volatile("" : "+r"(v[i]), "+r"(v[i + 1]), "+r"(v[i + 2]), "+r"(v[i + 3]), "+r"(v[i + 4]), "+r"(v[i + 5]), "+r"(v[i + 6]),
                          "+r"(v[i + 7]));
}

// Shared-memory descriptor of an MN-major operand tile written by TMA with SWIZZLE_128B: rows = K index (128 B each, the 64
// MN elements contiguous), 8-row / 1024 B swizzle atoms.  Canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units:
// SBO = distance between 8-row K groups (1024 B); LBO (distance between 64-element MN atoms) is unused for N = 64.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>(1u) << 16;
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16_bmn(uint32_t fmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// barrier slots
constexpr int kMaxQ = 4, kMaxKv = 6;
enum { kKvFull = 0, kKvEmpty = 6, kQFull = 12, kQEmpty = 16, kSFull = 20, kPFull = 22, kOFull = 24, kSEmpty = 26, kNumBars = 28 };

template <typename T, bool CAUSAL, bool EXTRA>
__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv_tail,
                    const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_o_tail,
                    const AttnTcParams p) {
    using H = Half16<T>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_o = smem;                          // 2 x [128 x 64] output staging tiles (one per softmax warpgroup)
    uint8_t* s_q = s_o + 2 * kTileQBytes;         // nq x [128 x 64] query tiles
    uint8_t* s_kv = s_q + p.nq * kTileQBytes;     // nkv stages x (K rows | V rows)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_kv + 2 * p.nkv * p.kv_stage_bytes);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = p.heads * kHd;
    const int L = p.L;
    const int q_tiles = p.q_tiles;
    const int n_items = (p.items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int n_tiles = n_items * q_tiles;

    // K / V rows beyond L are never written by TMA: zero the ring once (masked probabilities multiply finite values)
    for (int i = tid; i < (2 * p.nkv * p.kv_stage_bytes) / 16; i += kThreads) reinterpret_cast<uint4*>(s_kv)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < kMaxKv; ++s) {
            mbar_init(bars + kKvFull + s, 1);
            mbar_init(bars + kKvEmpty + s, EXTRA ? 1 + 128 * q_tiles : 1);
        }
        for (int s = 0; s < kMaxQ; ++s) {
            mbar_init(bars + kQFull + s, 1);
            mbar_init(bars + kQEmpty + s, EXTRA ? 128 : 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bars + kSFull + s, 1);
            mbar_init(bars + kPFull + s, 128);
            mbar_init(bars + kOFull + s, 1);
            mbar_init(bars + kSEmpty + s, 128);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv_tail);
        tma_prefetch_desc(&tmap_o);
        tma_prefetch_desc(&tmap_o_tail);
    }
    if (warp == 9) {
        tmem_alloc(tmem_ptr_smem, kTmemCols);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    pdl_launch_dependents();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int full_kv_boxes = L / 128;
    const int kv_tail = L - full_kv_boxes * 128;
    const bool tail_m64 = p.tail_rows <= 64;

    // 168 registers per thread at launch; the producer warpgroup hands 128 x 112 of them to the two softmax warpgroups
    // (each role branch issues its own setmaxnreg so that ptxas budgets the branch accordingly)
    if (warp == 8) {
        // ===================== TMA producer (one thread) =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (lane == 0) {
            int t = 0;
            for (int il = 0; il < n_items; ++il) {
                const int it = static_cast<int>(blockIdx.x) + il * static_cast<int>(gridDim.x);
                const int b = it / p.heads, h = it - b * p.heads;
                const int row0 = b * L;
                const int st = il % p.nkv;
                uint8_t* sk = s_kv + st * 2 * p.kv_stage_bytes;
                uint8_t* sv = sk + p.kv_stage_bytes;
                uint64_t* kv_full = bars + kKvFull + st;
                mbar_wait(bars + kKvEmpty + st, ((il / p.nkv) & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full, static_cast<uint32_t>(2 * L * 128));
                for (int bx = 0; bx < full_kv_boxes; ++bx) {
                    tma_load_2d(&tmap_q, kv_full, sk + bx * kTileQBytes, W + h * kHd, row0 + bx * 128, kCacheHintEvictFirst);
                    tma_load_2d(&tmap_q, kv_full, sv + bx * kTileQBytes, 2 * W + h * kHd, row0 + bx * 128, kCacheHintEvictFirst);
                }
                if (kv_tail > 0) {
                    tma_load_2d(&tmap_kv_tail, kv_full, sk + full_kv_boxes * kTileQBytes, W + h * kHd, row0 + full_kv_boxes * 128,
                                kCacheHintEvictFirst);
                    tma_load_2d(&tmap_kv_tail, kv_full, sv + full_kv_boxes * kTileQBytes, 2 * W + h * kHd, row0 + full_kv_boxes * 128,
                                kCacheHintEvictFirst);
                }
                for (int qt = 0; qt < q_tiles; ++qt, ++t) {
                    const int qs = t % p.nq;
                    mbar_wait(bars + kQEmpty + qs, ((t / p.nq) & 1) ^ 1);
                    mbar_arrive_expect_tx(bars + kQFull + qs, kTileQBytes);
                    tma_load_2d(&tmap_q, bars + kQFull + qs, s_q + qs * kTileQBytes, h * kHd, row0 + qt * kQT, kCacheHintEvictFirst);
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ===================== MMA issue (one thread) =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (lane == 0) {
            auto issue_pv = [&](int t) {
                const int g = t & 1;
                const int il = t / q_tiles, qt = t - il * q_tiles;
                const int st = il % p.nkv;
                const bool m64 = tail_m64 && qt == q_tiles - 1;
                const int ncols = CAUSAL ? min(p.n_cols, kQT * (qt + 1)) : p.n_cols;
                const uint32_t idesc_o = make_idesc_f16_bmn(H::kUmmaFormat, m64 ? 64 : 128, kHd);
                const uint64_t desc_v = make_sw128_mnmajor_desc(smem_u32(s_kv + st * 2 * p.kv_stage_bytes + p.kv_stage_bytes));
                const uint32_t tm = tmem_base + static_cast<uint32_t>(g * 256);
                mbar_wait(bars + kPFull + g, (t >> 1) & 1);
                tc_fence_after();
                B2C_STAMP(2, t);
                const int ksteps = ncols / 16;
                for (int j = 0; j < ksteps; ++j)
                    umma_f16_ts(tm + 128, tm + 8 * j, desc_v + static_cast<uint64_t>(j) * ((16u * 128u) >> 4), idesc_o, j != 0);
                umma_commit(bars + kOFull + g);
                if (qt == q_tiles - 1) umma_commit(bars + kKvEmpty + st);
                B2C_STAMP(7, t);
            };
            for (int t = 0; t < n_tiles; ++t) {
                const int g = t & 1;
                const int il = t / q_tiles, qt = t - il * q_tiles;
                const int st = il % p.nkv, qs = t % p.nq;
                const bool m64 = tail_m64 && qt == q_tiles - 1;
                const int ncols = CAUSAL ? min(p.n_cols, kQT * (qt + 1)) : p.n_cols;
                const uint32_t idesc_s = make_idesc_f16(H::kUmmaFormat, m64 ? 64 : 128, static_cast<uint32_t>(ncols));
                const uint64_t desc_q = make_sw128_kmajor_desc(smem_u32(s_q + qs * kTileQBytes));
                const uint64_t desc_k = make_sw128_kmajor_desc(smem_u32(s_kv + st * 2 * p.kv_stage_bytes));
                const uint32_t tm = tmem_base + static_cast<uint32_t>(g * 256);
                mbar_wait(bars + kSEmpty + g, ((t >> 1) & 1) ^ 1);      // O of tile t-2 has been drained from this TMEM half
                if (qt == 0) mbar_wait(bars + kKvFull + st, (il / p.nkv) & 1);
                mbar_wait(bars + kQFull + qs, (t / p.nq) & 1);
                tc_fence_after();
                B2C_STAMP(0, t);
#pragma unroll
                for (int k = 0; k < kHd / 16; ++k) umma_f16(tm, desc_q + 2 * k, desc_k + 2 * k, idesc_s, k != 0);
                umma_commit(bars + kSFull + g);
                if (!EXTRA) umma_commit(bars + kQEmpty + qs);
                if (t > 0) issue_pv(t - 1);
            }
            if (n_tiles > 0) issue_pv(n_tiles - 1);
        }
        __syncwarp();
    } else if (warp < 8) {
        // ===================== softmax + output (thread = query row; warpgroup g owns tiles t = g, g+2, ...) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int g = warp >> 2, wq = warp & 3;
        const int gtid = tid & 127;
        const uint32_t tm = tmem_base + static_cast<uint32_t>(g * 256) + (static_cast<uint32_t>(wq * 32) << 16);
        const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
        const uint32_t so = smem_u32(s_o + g * kTileQBytes);
        uint64_t* s_full = bars + kSFull + g;
        uint64_t* p_full = bars + kPFull + g;
        uint64_t* o_full = bars + kOFull + g;
        uint64_t* s_empty = bars + kSEmpty + g;
        for (int t = g; t < n_tiles; t += 2) {
            const uint32_t ph = (t >> 1) & 1;
            const int il = t / q_tiles, qt = t - il * q_tiles;
            const int it = static_cast<int>(blockIdx.x) + il * static_cast<int>(gridDim.x);
            const int b = it / p.heads, h = it - b * p.heads;
            const int st = il % p.nkv, qs = t % p.nq;
            const bool last = qt == q_tiles - 1;
            const bool m64 = tail_m64 && last;
            const int vr = last ? p.tail_rows : kQT;                      // valid rows of this tile
            const int r = m64 ? wq * 16 + lane : wq * 32 + lane;          // row inside the tile held by this lane
            const bool warp_live = (m64 ? wq * 16 : wq * 32) < vr;
            const int qrow = qt * kQT + r;                                // row inside the sequence
            const int ncols = CAUSAL ? min(p.n_cols, kQT * (qt + 1)) : p.n_cols;
            const int nch = (ncols + 31) >> 5;
            const uint32_t skv = smem_u32(s_kv + st * 2 * p.kv_stage_bytes);

            mbar_wait(s_full, ph);
            tc_fence_after();
            if (gtid == 0) B2C_STAMP(1, t);
            float sum = 0.f, s_x = 0.f, p_x = 0.f;
            if (warp_live) {
                // ---- pass 1: row maximum over the valid columns (loads kept one chunk ahead of the arithmetic)
                float mx = -INFINITY;
                auto max_chunk = [&](const uint32_t (&v)[32], int ch) {
                    const int c0 = ch * 32;
                    if (c0 + 32 > L || (CAUSAL && c0 + 32 > qrow + 1)) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = c0 + j;
                            const float s = (col >= L || (CAUSAL && col > qrow)) ? -INFINITY : __uint_as_float(v[j]);
                            mx = fmaxf(mx, s);
                        }
                    } else {
                        float m0 = mx, m1 = __uint_as_float(v[0]);   // two FMNMX3 chains
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            m0 = fmax3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                            m1 = fmax3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        }
                        mx = fmaxf(m0, m1);
                    }
                };
                {
                    // NOT unrolled over the chunks (the unrolled form overflowed the instruction cache: stall_no_inst), and the
                    // next chunk's tcgen05.ld is not kept in flight under this chunk's arithmetic: measured slower (113 vs 101 us
                    // at L = 197), the second softmax warpgroup on the same scheduler already fills the load latency.
                    uint32_t va[32];
#pragma unroll 1
                    for (int ch = 0; ch < nch; ++ch) {
                        tmem_ld_32x32(tm + ch * 32, va);
                        tmem_ld_wait_on(va);
                        max_chunk(va, ch);
                    }
                }
                if (EXTRA) {
                    // key 256 on the CUDA cores: s_x = q_row . k_256 (row 256 of the K stage: 256 % 8 == 0, so it is not swizzled)
                    const uint32_t qa = smem_u32(s_q + qs * kTileQBytes) + static_cast<uint32_t>(r) * 128, rx = static_cast<uint32_t>(r & 7);
                    const uint32_t ka = skv + 256u * 128u;
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (uint32_t ck = 0; ck < 8; ++ck) {
                        const uint4 qv = lds128(qa + ((ck ^ rx) << 4));
                        const uint4 kv = lds128(ka + (ck << 4));
                        const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w}, kw[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 qf = H::unpack(qw[e]), kf = H::unpack(kw[e]);
                            a0 = fmaf(qf.x, kf.x, a0);
                            a1 = fmaf(qf.y, kf.y, a1);
                        }
                    }
                    s_x = (CAUSAL && 256 > qrow) ? -INFINITY : a0 + a1;
                    mx = fmaxf(mx, s_x);
                }
                if (EXTRA) mbar_arrive(bars + kQEmpty + qs);
                // column 0 is valid for every row (causal: col 0 <= row), so mx is finite unless the scores themselves are not
                const float nm = -mx * c;
                // ---- pass 2: p = 2^(s*c - m*c), row sum, P (16-bit) over the S columns this thread has already consumed
                float sum0 = 0.f, sum1 = 0.f;
                uint64_t sum2a = 0ull, sum2b = 0ull;   // packed fp32x2 partial sums (FADD2)
                const uint64_t c2 = pack_f2(c, c), nm2 = pack_f2(nm, nm);
                auto exp_chunk = [&](const uint32_t (&v)[32], int ch) {
                    const int c0 = ch * 32;
                    uint32_t pk[16];
                    if (c0 + 32 > L || (CAUSAL && c0 + 32 > qrow + 1)) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float p0 = ex2_fast(fmaf(__uint_as_float(v[2 * j]), c, nm));
                            float p1 = ex2_fast(fmaf(__uint_as_float(v[2 * j + 1]), c, nm));
                            const int col = c0 + 2 * j;
                            if (col >= L || (CAUSAL && col > qrow)) p0 = 0.f;
                            if (col + 1 >= L || (CAUSAL && col + 1 > qrow)) p1 = 0.f;
                            sum0 += p0;
                            sum1 += p1;
                            pk[j] = H::pack(p0, p1);
                        }
                    } else {
                        // FFMA2 for the exponent argument and FADD2 for the row sum: 2.5 issue slots per element instead of 3.5
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float x0, x1;
                            unpack_f2(fma_f2(pack_f2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), c2, nm2), x0, x1);
                            const float p0 = ex2_fast(x0), p1 = ex2_fast(x1);
                            if (j & 1) sum2b = add_f2(sum2b, pack_f2(p0, p1));
                            else sum2a = add_f2(sum2a, pack_f2(p0, p1));
                            pk[j] = H::pack(p0, p1);
                        }
                    }
                    attn_tmem_st_32x16(tm + ch * 16, pk);
                };
                {
                    uint32_t va[32];
#pragma unroll 1
                    for (int ch = 0; ch < nch; ++ch) {
                        tmem_ld_32x32(tm + ch * 32, va);
                        tmem_ld_wait_on(va);
                        exp_chunk(va, ch);
                    }
                }
                {
                    float s0, s1;
                    unpack_f2(add_f2(sum2a, sum2b), s0, s1);
                    sum = (sum0 + sum1) + (s0 + s1);
                }
                if (EXTRA) {
                    p_x = ex2_fast(fmaf(s_x, c, nm));   // 2^(-inf) = 0 for the causally masked case
                    sum += p_x;
                }
                tmem_st_wait();
            } else if (EXTRA) {
                mbar_arrive(bars + kQEmpty + qs);
            }
            tc_fence_before();
            if (gtid == 0) B2C_STAMP(3, t);
            mbar_arrive(p_full);

            // ---- O: TMEM -> registers -> * 1/sum -> swizzled staging -> TMA store
            mbar_wait(o_full, ph);
            tc_fence_after();
            if (gtid == 0) B2C_STAMP(4, t);
            uint32_t o0[32], o1[32];
            if (warp_live) {
                tmem_ld_32x32(tm + 128, o0);
                tmem_ld_32x32(tm + 160, o1);
                tmem_ld_wait_on(o0);
                tmem_ld_wait_on(o1);
            }
            tc_fence_before();
            mbar_arrive(s_empty);
            if (gtid == 0) B2C_STAMP(5, t);
            if (gtid == 0) tma_store_wait_read<0>();   // this warpgroup's previous store has released the staging tile
            asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
            if (warp_live && (!m64 || lane < 16)) {
                const float inv = 1.0f / sum;
                const uint32_t row_off = static_cast<uint32_t>(r) * 128, rx = static_cast<uint32_t>(r & 7);
#pragma unroll
                for (int gq = 0; gq < 8; ++gq) {
                    const uint32_t* src = gq < 4 ? &o0[gq * 8] : &o1[(gq - 4) * 8];
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(src[e]);
                    if (EXTRA) {
                        const uint4 vv = lds128(skv + static_cast<uint32_t>(p.kv_stage_bytes) + 256u * 128u + (static_cast<uint32_t>(gq) << 4));
                        const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 vf = H::unpack(vw[e]);
                            f[2 * e] = fmaf(p_x, vf.x, f[2 * e]);
                            f[2 * e + 1] = fmaf(p_x, vf.y, f[2 * e + 1]);
                        }
                    }
                    uint4 w;
                    w.x = H::pack(f[0] * inv, f[1] * inv);
                    w.y = H::pack(f[2] * inv, f[3] * inv);
                    w.z = H::pack(f[4] * inv, f[5] * inv);
                    w.w = H::pack(f[6] * inv, f[7] * inv);
                    const uint32_t addr = so + row_off + ((static_cast<uint32_t>(gq) ^ rx) << 4);
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
                }
            }
            if (EXTRA) mbar_arrive(bars + kKvEmpty + st);   // row 256 of K and V is no longer read by this thread
            fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
            if (gtid == 0) {
                const bool tail = last && p.tail_rows != kQT;
                tma_store_2d(tail ? &tmap_o_tail : &tmap_o, s_o + g * kTileQBytes, h * kHd, b * L + qt * kQT);
                tma_store_commit();
                B2C_STAMP(6, t);
            }
        }
        if (gtid == 0) tma_store_wait_all<0>();
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <typename T, bool CAUSAL, bool EXTRA>
cudaError_t launch_tc(const cudaLaunchConfig_t& cfg, const CUtensorMap& tq, const CUtensorMap& tkvt, const CUtensorMap& to,
                      const CUtensorMap& tot, const AttnTcParams& p) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(attention_tc_kernel<T, CAUSAL, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    });
    if (attr_err != cudaSuccess) return attr_err;
    return cudaLaunchKernelEx(&cfg, attention_tc_kernel<T, CAUSAL, EXTRA>, tq, tkvt, to, tot, p);
}

}  // namespace

static long long* g_attn_dbg = nullptr;
void attention_tc_set_debug(long long* buf) { g_attn_dbg = buf; }

// qkv [batch*L, 3*heads*64] -> out [batch*L, heads*64]; 16-bit dtypes, 64 < L <= 257.  Returns 1 when the shape is outside
// this kernel's range (the caller then uses the generic path), 0 on success, < 0 / CUDA code on error.
int attention_tc(int dtype, const void* qkv, void* out, int batch, int L, int heads, int causal, cudaStream_t stream) {
    if (!(dtype == 1 || dtype == 2) || L <= 64 || L > 257) return 1;
    const int W = heads * kHd;
    const int64_t rows = static_cast<int64_t>(batch) * L;
    const bool extra = L == 257;
    AttnTcParams p;
    p.L = L;
    p.heads = heads;
    p.q_tiles = (L + kQT - 1) / kQT;
    const int64_t items = static_cast<int64_t>(batch) * heads;
    B2C_CHECK_ARG(items <= 0x7fffffff / 4, "attention: too many work items");
    p.items = static_cast<int>(items);
    p.tail_rows = L - (p.q_tiles - 1) * kQT;
    p.n_cols = extra ? 256 : (L + 15) / 16 * 16;
    p.kv_stage_bytes = (extra ? 264 : p.n_cols) * 128;
    p.dbg = g_attn_dbg;
    const int kv_tail = L % 128;

    const bool bf = dtype == 1;
    CUtensorMap tq, tkvt, to, tot;
    if (make_tmap_2d(&tq, bf, qkv, rows, 3 * W, 3 * W, kQT, kHd) != 0) return -1;
    if (make_tmap_2d(&tkvt, bf, qkv, rows, 3 * W, 3 * W, kv_tail > 0 ? kv_tail : 128, kHd) != 0) return -1;
    if (make_tmap_2d(&to, bf, out, rows, W, W, kQT, kHd) != 0) return -1;
    if (make_tmap_2d(&tot, bf, out, rows, W, W, p.tail_rows, kHd) != 0) return -1;

    // ring depths: the minimum is 2 + 2; query slots first (their prefetch distance is one tile), then K/V stages
    const int fixed_bytes = 2 * kTileQBytes + kNumBars * 8 + 16 + 1024;
    p.nq = 2;
    p.nkv = 2;
    auto total = [&](int nq, int nkv) { return fixed_bytes + nq * kTileQBytes + 2 * nkv * p.kv_stage_bytes; };
    while (p.nq < kMaxQ && total(p.nq + 1, p.nkv) <= kMaxSmem) ++p.nq;
    while (p.nkv < kMaxKv && total(p.nq, p.nkv + 1) <= kMaxSmem) ++p.nkv;
    const int smem_bytes = total(p.nq, p.nkv);
    B2C_CHECK_ARG(smem_bytes <= kMaxSmem, "attention_tc: shared memory budget exceeded");
    const int grid = static_cast<int>(items < num_sms() ? items : num_sms());

    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t le;
    if (bf) {
        if (extra) le = causal ? launch_tc<__nv_bfloat16, true, true>(cfg, tq, tkvt, to, tot, p) : launch_tc<__nv_bfloat16, false, true>(cfg, tq, tkvt, to, tot, p);
        else le = causal ? launch_tc<__nv_bfloat16, true, false>(cfg, tq, tkvt, to, tot, p) : launch_tc<__nv_bfloat16, false, false>(cfg, tq, tkvt, to, tot, p);
    } else {
        if (extra) le = causal ? launch_tc<__half, true, true>(cfg, tq, tkvt, to, tot, p) : launch_tc<__half, false, true>(cfg, tq, tkvt, to, tot, p);
        else le = causal ? launch_tc<__half, true, false>(cfg, tq, tkvt, to, tot, p) : launch_tc<__half, false, false>(cfg, tq, tkvt, to, tot, p);
    }
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(attention_tc_kernel)");
    B2C_LAUNCH_CHECK("attention_tc_kernel");
    return 0;
}

}  // namespace b200clip
