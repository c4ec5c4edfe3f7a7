// Fused flash-style multi-head attention for the CLIP towers (head_dim = 64, L in {50,77,197,257,...}).
//
// Replaces the SDPA core of nn.MultiheadAttention as used by ResidualAttentionBlock.attention
// (deps/open_clip/src/open_clip/transformer.py:224,238-251): softmax(q k^T / sqrt(64) + mask) v with the
// text tower's strict upper-triangular -inf mask (transformer.py:751-757) as the `causal` variant.
//
// 16-bit path: one CTA = 64 query rows of one (batch, head); 4 warps x 16 rows.  Q/K/V tiles are staged in
// XOR-swizzled shared memory with cp.async, scores and the running (max, sum) stay in registers, the
// row reductions are quad shuffles, P is re-used straight from the score accumulators as the A operand
// of the P·V product (no shared-memory round trip).  Tensor work is mma.sync m16n8k16 with fp32
// accumulation: at L <= 257 the kernel is bound by the qkv read / out write, not by the tensor pipe.
// fp32 path (parity mode): SIMT, one warp per query row, fp32 everywhere.
#include "common.cuh"
#include "internal.h"
#include "tmap.h"

#include <cstdlib>
#include <mutex>

namespace b200clip {

int attention_tc(int dtype, const void* qkv, void* out, int batch, int L, int heads, int causal, cudaStream_t stream);

namespace {

constexpr int kHeadDim = 64;
constexpr int kBlockQ = 64;
constexpr int kBlockKV = 64;

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

template <typename T> __device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of (row, 16-byte chunk) inside a [64 rows][128 B] tile with the 8-chunk XOR swizzle
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// Stage rows [row0, row0+64) of one 64-wide column block of the packed qkv matrix into a swizzled tile.
template <typename T>
__device__ __forceinline__ void load_tile(uint32_t smem_tile, const T* base, int64_t ld, int row0, int rows_total, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 128;
        const int r = idx >> 3, ch = idx & 7;
        const bool ok = row0 + r < rows_total;
        const T* src = base + static_cast<int64_t>(ok ? row0 + r : 0) * ld + ch * 8;
        cp_async16(smem_tile + swz(r, ch), src, ok);
    }
}

template <typename T, bool CAUSAL>
__global__ void __launch_bounds__(128)
attention_mma_kernel(const T* __restrict__ qkv, T* __restrict__ out, int L, int heads) {
    using H = Half16<T>;
    __shared__ __align__(128) uint8_t sQ[kBlockQ * 128];
    __shared__ __align__(128) uint8_t sK[2][kBlockKV * 128];
    __shared__ __align__(128) uint8_t sV[2][kBlockKV * 128];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.x;
    const int b = bh / heads, h = bh - b * heads;
    const int W = heads * kHeadDim;
    const int64_t ld = 3 * static_cast<int64_t>(W);
    const int q0 = blockIdx.y * kBlockQ;
    const T* qbase = qkv + static_cast<int64_t>(b) * L * ld + h * kHeadDim;
    const T* kbase = qbase + W;
    const T* vbase = qbase + 2 * W;

    const uint32_t sq = smem_u32(sQ);
    const uint32_t sk0 = smem_u32(sK[0]), sv0 = smem_u32(sV[0]);
    constexpr uint32_t kTileBytes = kBlockKV * 128;

    int kv_end = L;
    if (CAUSAL) kv_end = min(L, q0 + kBlockQ);
    const int nblk = (kv_end + kBlockKV - 1) / kBlockKV;

    load_tile<T>(sq, qbase, ld, q0, L, tid);
    load_tile<T>(sk0, kbase, ld, 0, L, tid);
    load_tile<T>(sv0, vbase, ld, 0, L, tid);
    cp_async_commit();

    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const float scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)

    uint32_t qf[4][4];
    const int qrow_a = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;  // ldmatrix row for the A fragments
    const int row_lo = q0 + warp * 16 + (lane >> 2);                      // query row of c0,c1 (c2,c3: +8)

    for (int blk = 0; blk < nblk; ++blk) {
        const int buf = blk & 1;
        if (blk + 1 < nblk) {
            load_tile<T>(sk0 + (buf ^ 1) * kTileBytes, kbase, ld, (blk + 1) * kBlockKV, L, tid);
            load_tile<T>(sv0 + (buf ^ 1) * kTileBytes, vbase, ld, (blk + 1) * kBlockKV, L, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (blk == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(sq + swz(qrow_a, kk * 2 + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
        }
        const uint32_t sk = sk0 + buf * kTileBytes, sv = sv0 + buf * kTileBytes;

        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {  // pairs of 8-wide kv blocks
                uint32_t b0, b1, b2, b3;
                const int krow = np * 16 + (lane & 7) + (lane >> 4) * 8;
                ldmatrix_x4(sk + swz(krow, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
                mma16816<T>(s[2 * np], qf[kk], b0, b1);
                mma16816<T>(s[2 * np + 1], qf[kk], b2, b3);
            }
        }

        // mask (only the block holding the sequence tail and, for the causal variant, the diagonal block need it)
        const int kv0 = blk * kBlockKV;
        const bool need_mask = kv0 + kBlockKV > L || (CAUSAL && kv0 + kBlockKV > q0);
        if (need_mask) {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col = kv0 + nb * 8 + (lane & 3) * 2 + (j & 1);
                    const int row = row_lo + (j >> 1) * 8;
                    if (col >= L || (CAUSAL && col > row)) s[nb][j] = -INFINITY;
                }
            }
        }
        // online softmax on the raw scores: p = 2^(s*c - m*c) is one FFMA + one MUFU per element
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
        }
        float corr[2], nm[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);  // finite: kv column 0 is never masked
            corr[r] = ex2_fast((m_run[r] - m_new) * scale_log2);
            m_run[r] = m_new;
            nm[r] = -m_new * scale_log2;
            l_run[r] *= corr[r];
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[4][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const float p0 = ex2_fast(fmaf(s[nb][0], scale_log2, nm[0]));
            const float p1 = ex2_fast(fmaf(s[nb][1], scale_log2, nm[0]));
            const float p2 = ex2_fast(fmaf(s[nb][2], scale_log2, nm[1]));
            const float p3 = ex2_fast(fmaf(s[nb][3], scale_log2, nm[1]));
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            // accumulator (row, 2 cols) pairs are exactly the A-fragment registers of the P·V product
            pf[nb >> 1][(nb & 1) * 2 + 0] = H::pack(p0, p1);
            pf[nb >> 1][(nb & 1) * 2 + 1] = H::pack(p2, p3);
        }
        l_run[0] += rs[0];
        l_run[1] += rs[1];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            o[nb][0] *= corr[0];
            o[nb][1] *= corr[0];
            o[nb][2] *= corr[1];
            o[nb][3] *= corr[1];
        }

        // O += P V  (k = kv index, n = head dim)
#pragma unroll
        for (int j = 0; j < 4; ++j) {      // 16-wide kv steps
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-wide d blocks
                uint32_t b0, b1, b2, b3;
                const int vrow = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldmatrix_x4_trans(sv + swz(vrow, dp * 2 + (lane >> 4)), b0, b1, b2, b3);
                mma16816<T>(o[2 * dp], pf[j], b0, b1);
                mma16816<T>(o[2 * dp + 1], pf[j], b2, b3);
            }
        }
        __syncthreads();  // everyone is done with buffer `buf` before the next prefetch overwrites it
    }

    // finalise: divide by the row sums (quad-reduced) and store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    T* obase = out + static_cast<int64_t>(b) * L * W + h * kHeadDim;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const int col = nb * 8 + (lane & 3) * 2;
        if (row_lo < L) *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(row_lo) * W + col) = H::pack(o[nb][0] * inv0, o[nb][1] * inv0);
        if (row_lo + 8 < L)
            *reinterpret_cast<uint32_t*>(obase + static_cast<int64_t>(row_lo + 8) * W + col) = H::pack(o[nb][2] * inv1, o[nb][3] * inv1);
    }
}

// ------------------------------- short-sequence fast path (L <= 64) -------------------------------
// ViT-B/32 (L = 50) and EOT-truncated text run here.  One (batch, head) item fits a single 64 x 64 score tile, so the kernel
// is a stream of small independent items and is bound by the qkv read / out write: persistent CTAs walk the items with a
// 2-stage TMA ring (three CTAs per SM) (warp 4 = producer: three [L x 64] boxes per item straight out of the packed qkv matrix, 128B-swizzled;
// warps 0-3 = consumers, 16 query rows each), so the loads of item i+1 are in flight while item i is computed, and the
// output tile leaves through a TMA store (full 128-byte rows instead of 4-byte scatters).
constexpr int kTileBytes = 64 * 128;
constexpr int short_smem_bytes(int stages) { return stages * 3 * kTileBytes + 2 * kTileBytes + 128 + 1024; }

// NKB = ceil(L / 8): number of 8-wide kv blocks that hold valid columns (compile-time so that every loop below unrolls
// into straight-line code; a run-time bound made the compiler fall back to jump tables and local-memory arrays)
template <typename T, bool CAUSAL, int NKB>
__global__ void __launch_bounds__(160, 3)
attention_short_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out, int L, int heads,
                       int items) {
    using H = Half16<T>;
    constexpr int kShortStages = 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;                                      // [stage][q|k|v][64 rows x 128 B]
    uint8_t* obuf = smem + kShortStages * 3 * kTileBytes;      // [2][64 x 128 B]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(obuf + 2 * kTileBytes);
    uint64_t* empty_bar = full_bar + kShortStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = heads * kHeadDim;

    // rows [L, 64) of the tiles are never written by TMA: zero them once so masked lanes multiply finite values
    for (int i = tid; i < (kShortStages * 3 + 2) * kTileBytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < kShortStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 4);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmap_qkv);
        tma_prefetch_desc(&tmap_out);
    }
    fence_proxy_async();  // the zero fill (generic proxy) is ordered before the TMA writes (async proxy)
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();  // the qkv matrix is the previous kernel's output; everything above overlapped its tail

    const uint32_t tile_tx = static_cast<uint32_t>(L) * 128u;
    if (warp == 4) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int b = it / heads, h = it - b * heads;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full_bar[stage], 3 * tile_tx);
                uint8_t* dst = ring + stage * 3 * kTileBytes;
                tma_load_2d(&tmap_qkv, &full_bar[stage], dst, h * kHeadDim, b * L, kCacheHintEvictFirst);
                tma_load_2d(&tmap_qkv, &full_bar[stage], dst + kTileBytes, W + h * kHeadDim, b * L, kCacheHintEvictFirst);
                tma_load_2d(&tmap_qkv, &full_bar[stage], dst + 2 * kTileBytes, 2 * W + h * kHeadDim, b * L, kCacheHintEvictFirst);
                if (++stage == kShortStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }

    // ---- consumers: warp w owns query rows [16w, 16w+16) ----
    const float scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const int qrow_a = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int row_lo = warp * 16 + (lane >> 2);
    constexpr int nkb = NKB;             // 8-wide kv blocks that contain valid columns
    constexpr int nkk = (NKB + 1) / 2;   // 16-deep kv steps of the P.V product
    const bool warp_active = warp * 16 < L;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t oc = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++oc) {
        const int b = it / heads, h = it - b * heads;
        mbar_wait(&full_bar[stage], phase);
        const uint32_t sq = smem_u32(ring + stage * 3 * kTileBytes);
        const uint32_t sk = sq + kTileBytes, sv = sq + 2 * kTileBytes;
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
        float inv0 = 0.f, inv1 = 0.f;
        if (warp_active) {
            uint32_t qf[4][4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(sq + swz(qrow_a, kk * 2 + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
            float s[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    if (np * 2 < nkb) {
                        uint32_t b0, b1, b2, b3;
                        const int krow = np * 16 + (lane & 7) + (lane >> 4) * 8;
                        ldmatrix_x4(sk + swz(krow, kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
                        mma16816<T>(s[2 * np], qf[kk], b0, b1);
                        if (np * 2 + 1 < nkb) mma16816<T>(s[2 * np + 1], qf[kk], b2, b3);
                    }
                }
            }
            // mask: only the last valid 8-wide block can contain columns >= L; the causal variant masks per element.
            // Blocks >= nkb were never computed and are treated as -inf (probability 0) below.
            if constexpr (CAUSAL) {
#pragma unroll
                for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col = nb * 8 + (lane & 3) * 2 + (j & 1);
                        const int row = row_lo + (j >> 1) * 8;
                        if (col >= L || col > row) s[nb][j] = -INFINITY;
                    }
                }
            } else {
#pragma unroll
                for (int nb = 0; nb < 8; ++nb) {
                    if (nb == nkb - 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (nb * 8 + (lane & 3) * 2 + (j & 1) >= L) s[nb][j] = -INFINITY;
                    }
                }
            }
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                if (nb < nkb) {
                    mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
                    mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            }
            // p = 2^(s*c - max*c): one FFMA + one MUFU per element (column 0 is never masked, so max is finite)
            const float nm0 = -mx[0] * scale_log2, nm1 = -mx[1] * scale_log2;
            float rs[2] = {0.f, 0.f};
            uint32_t pf[4][4];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
                if (nb < nkb) {
                    p0 = ex2_fast(fmaf(s[nb][0], scale_log2, nm0));
                    p1 = ex2_fast(fmaf(s[nb][1], scale_log2, nm0));
                    p2 = ex2_fast(fmaf(s[nb][2], scale_log2, nm1));
                    p3 = ex2_fast(fmaf(s[nb][3], scale_log2, nm1));
                    rs[0] += p0 + p1;
                    rs[1] += p2 + p3;
                }
                pf[nb >> 1][(nb & 1) * 2 + 0] = H::pack(p0, p1);
                pf[nb >> 1][(nb & 1) * 2 + 1] = H::pack(p2, p3);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
                rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
            }
            inv0 = 1.f / rs[0];
            inv1 = 1.f / rs[1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nkk) {
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {
                        uint32_t b0, b1, b2, b3;
                        const int vrow = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                        ldmatrix_x4_trans(sv + swz(vrow, dp * 2 + (lane >> 4)), b0, b1, b2, b3);
                        mma16816<T>(o[2 * dp], pf[j], b0, b1);
                        mma16816<T>(o[2 * dp + 1], pf[j], b2, b3);
                    }
                }
            }
        }
        // all shared-memory reads of this stage are done: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == kShortStages) {
            stage = 0;
            phase ^= 1;
        }

        // ---- output tile: registers -> swizzled staging -> TMA store ----
        const uint32_t ob = oc & 1;
        if (tid == 0) tma_store_wait_read<1>();  // the store issued two items ago has released this buffer
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint32_t so = smem_u32(obuf + ob * kTileBytes);
        if (warp_active) {
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const uint32_t off = static_cast<uint32_t>(lane & 3) * 4;
                const uint32_t a0 = so + swz(row_lo, nb) + off;
                const uint32_t a1 = so + swz(row_lo + 8, nb) + off;
                const uint32_t w0 = H::pack(o[nb][0] * inv0, o[nb][1] * inv0);
                const uint32_t w1 = H::pack(o[nb][2] * inv1, o[nb][3] * inv1);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(a0), "r"(w0) : "memory");
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(a1), "r"(w1) : "memory");
            }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid == 0) {
            tma_store_2d(&tmap_out, obuf + ob * kTileBytes, h * kHeadDim, b * L);
            tma_store_commit();
        }
    }
    if (tid == 0) tma_store_wait_all<0>();
}

// ------------------------------- fp32 parity path -------------------------------
// One CTA per (batch, head); K and V of the head live in shared memory (row pitch 65 floats: conflict-free
// for both the per-lane-row dot products and the per-lane-column P·V pass); one warp per query row.
__global__ void __launch_bounds__(256)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int L, int heads, int causal) {
    extern __shared__ float smem_f[];
    constexpr int P = kHeadDim + 1;
    float* sK = smem_f;
    float* sV = sK + static_cast<size_t>(L) * P;
    float* sW = sV + static_cast<size_t>(L) * P;  // per warp: q[64] + p[L]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int bh = blockIdx.x;
    const int b = bh / heads, h = bh - b * heads;
    const int W = heads * kHeadDim;
    const int64_t ld = 3 * static_cast<int64_t>(W);
    const float* base = qkv + static_cast<int64_t>(b) * L * ld + h * kHeadDim;
    for (int i = threadIdx.x; i < L * kHeadDim; i += blockDim.x) {
        const int r = i >> 6, c = i & 63;
        sK[r * P + c] = base[static_cast<int64_t>(r) * ld + W + c];
        sV[r * P + c] = base[static_cast<int64_t>(r) * ld + 2 * W + c];
    }
    __syncthreads();
    float* sq = sW + static_cast<size_t>(warp) * (kHeadDim + L);
    float* sp = sq + kHeadDim;
    for (int q = warp; q < L; q += nwarps) {
        sq[lane] = base[static_cast<int64_t>(q) * ld + lane];
        sq[lane + 32] = base[static_cast<int64_t>(q) * ld + lane + 32];
        __syncwarp();
        const int kv_end = causal ? q + 1 : L;
        float mx = -INFINITY;
        for (int kv = lane; kv < kv_end; kv += 32) {
            float acc = 0.f;
#pragma unroll 16
            for (int d = 0; d < kHeadDim; ++d) acc = fmaf(sq[d], sK[kv * P + d], acc);
            acc *= 0.125f;
            sp[kv] = acc;
            mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int kv = lane; kv < kv_end; kv += 32) {
            const float e = expf(sp[kv] - mx);
            sp[kv] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int kv = 0; kv < kv_end; ++kv) {
            const float p = sp[kv];
            o0 = fmaf(p, sV[kv * P + lane], o0);
            o1 = fmaf(p, sV[kv * P + lane + 32], o1);
        }
        const float inv = 1.f / sum;
        float* orow = out + (static_cast<int64_t>(b) * L + q) * W + h * kHeadDim;
        orow[lane] = o0 * inv;
        orow[lane + 32] = o1 * inv;
        __syncwarp();
    }
}

template <typename T, bool CAUSAL, int NKB>
int launch_short_one(int grid, const CUtensorMap& tq, const CUtensorMap& to, int L, int heads, int items, cudaStream_t stream) {
    auto kern = attention_short_kernel<T, CAUSAL, NKB>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, short_smem_bytes(2)); });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(attention_short smem)");
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(160);
    cfg.dynamicSmemBytes = short_smem_bytes(2);
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tq, to, L, heads, items);
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(attention_short_kernel)");
    return 0;
}
template <typename T, bool CAUSAL>
int launch_short_nkb(int nkb, int grid, const CUtensorMap& tq, const CUtensorMap& to, int L, int heads, int items, cudaStream_t s) {
    switch (nkb) {
        case 1: return launch_short_one<T, CAUSAL, 1>(grid, tq, to, L, heads, items, s);
        case 2: return launch_short_one<T, CAUSAL, 2>(grid, tq, to, L, heads, items, s);
        case 3: return launch_short_one<T, CAUSAL, 3>(grid, tq, to, L, heads, items, s);
        case 4: return launch_short_one<T, CAUSAL, 4>(grid, tq, to, L, heads, items, s);
        case 5: return launch_short_one<T, CAUSAL, 5>(grid, tq, to, L, heads, items, s);
        case 6: return launch_short_one<T, CAUSAL, 6>(grid, tq, to, L, heads, items, s);
        case 7: return launch_short_one<T, CAUSAL, 7>(grid, tq, to, L, heads, items, s);
        case 8: return launch_short_one<T, CAUSAL, 8>(grid, tq, to, L, heads, items, s);
    }
    set_last_error("attention: bad kv block count %d", nkb);
    return -1;
}
int launch_short(bool is_bf16, bool causal, int nkb, int grid, const CUtensorMap& tq, const CUtensorMap& to, int L, int heads, int items,
                 cudaStream_t s) {
    if (is_bf16) return causal ? launch_short_nkb<__nv_bfloat16, true>(nkb, grid, tq, to, L, heads, items, s)
                               : launch_short_nkb<__nv_bfloat16, false>(nkb, grid, tq, to, L, heads, items, s);
    return causal ? launch_short_nkb<__half, true>(nkb, grid, tq, to, L, heads, items, s)
                  : launch_short_nkb<__half, false>(nkb, grid, tq, to, L, heads, items, s);
}

// B200CLIP_ATTN_GENERIC=1 routes short sequences to the generic flash kernel too (A/B measurements)
bool attention_force_generic() {
    static const bool v = [] {
        const char* e = getenv("B200CLIP_ATTN_GENERIC");
        return e != nullptr && e[0] == '1';
    }();
    return v;
}

}  // namespace

int attention(int dtype, const void* qkv, void* out, int batch, int seq_len, int heads, int causal, cudaStream_t stream) {
    B2C_CHECK_ARG(batch > 0 && seq_len > 0 && heads > 0, "attention: bad shape batch=%d L=%d heads=%d", batch, seq_len, heads);
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
                  "attention: pointers must be 16-byte aligned");
    const int64_t bh = static_cast<int64_t>(batch) * heads;
    if (dtype == 0) {
        B2C_CHECK_ARG(bh <= 0x7fffffff, "attention: batch*heads too large");
        const int nwarps = 8;
        const size_t smem = (2 * static_cast<size_t>(seq_len) * (kHeadDim + 1) + nwarps * (kHeadDim + seq_len)) * sizeof(float);
        B2C_CHECK_ARG(smem <= 227 * 1024, "attention(fp32): sequence length %d too long for the shared-memory path", seq_len);
        static size_t configured = 0;
        if (smem > 48 * 1024 && smem > configured) {
            B2C_CUDA(cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            configured = 227 * 1024;
        }
        attention_f32_kernel<<<static_cast<unsigned>(bh), nwarps * 32, smem, stream>>>(static_cast<const float*>(qkv),
                                                                                        static_cast<float*>(out), seq_len, heads, causal);
        B2C_LAUNCH_CHECK("attention_f32_kernel");
        return 0;
    }
    B2C_CHECK_ARG(bh <= 0x7fffffff, "attention: batch*heads too large");
    if (seq_len <= 64 && (dtype == 1 || dtype == 2) && !attention_force_generic()) {
        const int W = heads * kHeadDim;
        const int64_t rows = static_cast<int64_t>(batch) * seq_len;
        CUtensorMap tq, to;
        if (make_tmap_2d(&tq, dtype == 1, qkv, rows, 3 * W, 3 * W, seq_len, kHeadDim) != 0) return -1;
        if (make_tmap_2d(&to, dtype == 1, out, rows, W, W, seq_len, kHeadDim) != 0) return -1;
        const int64_t max_ctas = 3 * static_cast<int64_t>(num_sms());
        const int grid_s = static_cast<int>(bh < max_ctas ? bh : max_ctas);
        const int rc = launch_short(dtype == 1, causal != 0, (seq_len + 7) / 8, grid_s, tq, to, seq_len, heads, static_cast<int>(bh), stream);
        if (rc != 0) return rc;
        B2C_LAUNCH_CHECK("attention_short_kernel");
        return 0;
    }
    if (!attention_force_generic()) {
        // 64 < L <= 288: tcgen05 kernel (attention_tc.cu); returns 1 when the shape is outside its range
        const int rc = attention_tc(dtype, qkv, out, batch, seq_len, heads, causal, stream);
        if (rc != 1) return rc;
    }
    dim3 grid(static_cast<unsigned>(bh), (seq_len + kBlockQ - 1) / kBlockQ);
    const __nv_bfloat16* qb = static_cast<const __nv_bfloat16*>(qkv);
    const __half* qh = static_cast<const __half*>(qkv);
    if (dtype == 1 && causal) attention_mma_kernel<__nv_bfloat16, true><<<grid, 128, 0, stream>>>(qb, static_cast<__nv_bfloat16*>(out), seq_len, heads);
    else if (dtype == 1) attention_mma_kernel<__nv_bfloat16, false><<<grid, 128, 0, stream>>>(qb, static_cast<__nv_bfloat16*>(out), seq_len, heads);
    else if (dtype == 2 && causal) attention_mma_kernel<__half, true><<<grid, 128, 0, stream>>>(qh, static_cast<__half*>(out), seq_len, heads);
    else if (dtype == 2) attention_mma_kernel<__half, false><<<grid, 128, 0, stream>>>(qh, static_cast<__half*>(out), seq_len, heads);
    else {
        set_last_error("attention: unknown dtype %d", dtype);
        return -1;
    }
    B2C_LAUNCH_CHECK("attention_mma_kernel");
    return 0;
}

}  // namespace b200clip
