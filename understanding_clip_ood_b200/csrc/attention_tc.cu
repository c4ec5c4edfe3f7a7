// tcgen05 attention for the longer sequences of the CLIP towers (64 < L <= 288: ViT-B/16 L=197, ViT-L/14 L=257, the
// untruncated text context L=77), head_dim 64.  Replaces the SDPA core of nn.MultiheadAttention as used by
// ResidualAttentionBlock.attention (deps/open_clip/src/open_clip/transformer.py:224,238-251), causal variant = the text
// tower's strict upper-triangular mask (transformer.py:751-757).
//
// One work item = 128 query rows of one (batch, head).  The whole key/value range fits one accumulator tile (L <= 288
// columns of TMEM), so there is no online-softmax rescaling:
//   warp 4 (one thread)  TMA: Q box [128 x 64], K and V boxes [L x 64] straight out of the packed qkv matrix (128B swizzle)
//                        tcgen05.mma  S[128 x Lpad] = Q K^T      (SS, accumulator in TMEM)
//                        tcgen05.mma  O[128 x 64]  = P V          (A = P read from TMEM, B = V as an MN-major smem operand:
//                                                                  no transpose of V anywhere)
//   warps 0-3            thread = query row: two passes over its S row with tcgen05.ld (row max, then p = 2^(s*c - m*c),
//                        row sum, P packed to 16-bit and written back with tcgen05.st over the S columns already consumed),
//                        later O * (1 / sum) -> swizzled staging tile -> TMA store.
// P aliases the first Lpad/2 columns of S, O the next 64, so an item needs Lpad TMEM columns (two CTAs per SM up to L = 256).
#include "common.cuh"
#include "internal.h"
#include "tmap.h"

#include <mutex>

namespace b200clip {

namespace {

constexpr int kHd = 64;
constexpr int kQT = 128;               // query rows per work item
constexpr int kTileQBytes = kQT * 128;

struct AttnTcParams {
    int L, Lpad, heads, q_tiles, items, tail_rows, tmem_cols;
};

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
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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

template <typename T, bool CAUSAL>
__global__ void __launch_bounds__(160)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_kv_tail, const __grid_constant__ CUtensorMap tmap_o,
                    const __grid_constant__ CUtensorMap tmap_o_tail, const AttnTcParams p) {
    using H = Half16<T>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kv_bytes = p.Lpad * 128;
    uint8_t* s_q = smem;
    uint8_t* s_k = s_q + kTileQBytes;
    uint8_t* s_v = s_k + kv_bytes;
    uint8_t* s_o = s_v + kv_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_o + kTileQBytes);
    uint64_t* bar_load = bars + 0;   // TMA bytes of Q, K, V
    uint64_t* bar_s = bars + 1;      // S = Q K^T complete (tcgen05.commit)
    uint64_t* bar_p = bars + 2;      // P written by all 128 softmax threads
    uint64_t* bar_o = bars + 3;      // O = P V complete (tcgen05.commit)
    uint64_t* bar_done = bars + 4;   // O read out of TMEM by all 128 threads: the accumulator columns may be overwritten
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = p.heads * kHd;
    const int L = p.L;

    // K / V rows [L, Lpad) are never written by TMA: zero them once (masked probabilities multiply finite values)
    for (int i = tid; i < (2 * kv_bytes) / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_k)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 128);
        mbar_init(bar_o, 1);
        mbar_init(bar_done, 128);
        fence_mbar_init();
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        tma_prefetch_desc(&tmap_o);
    }
    if (warp == 4) {
        tmem_alloc(tmem_ptr_smem, static_cast<uint32_t>(p.tmem_cols));
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    pdl_launch_dependents();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_s = tmem_base;                                   // S: columns [0, Lpad)
    const uint32_t tmem_p = tmem_base;                                   // P: columns [0, Lpad/2), over consumed S columns
    const uint32_t tmem_o = tmem_base + static_cast<uint32_t>(p.Lpad / 2);  // O: the next 64 columns
    const int full_kv_boxes = L / 128;
    const int kv_tail = L - full_kv_boxes * 128;

    if (warp == 4) {
        // ===================== TMA + MMA issue (one thread) =====================
        if (lane == 0) {
            const uint32_t n1 = static_cast<uint32_t>(p.Lpad > 256 ? 256 : p.Lpad);
            const uint32_t n2 = static_cast<uint32_t>(p.Lpad) - n1;
            const uint32_t idesc_s1 = make_idesc_f16(H::kUmmaFormat, kQT, n1);
            const uint32_t idesc_s2 = make_idesc_f16(H::kUmmaFormat, kQT, n2 > 0 ? n2 : 16);
            const uint32_t idesc_o = make_idesc_f16_bmn(H::kUmmaFormat, kQT, kHd);
            const uint64_t desc_q = make_sw128_kmajor_desc(smem_u32(s_q));
            const uint64_t desc_k = make_sw128_kmajor_desc(smem_u32(s_k));
            const uint64_t desc_v = make_sw128_mnmajor_desc(smem_u32(s_v));
            uint32_t it_n = 0;
            for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++it_n) {
                const uint32_t ph = it_n & 1;
                const int bh = it / p.q_tiles, qt = it - bh * p.q_tiles;
                const int b = bh / p.heads, h = bh - b * p.heads;
                const int row0 = b * L;
                // the previous item's MMAs have finished reading Q / K / V (bar_o) and its accumulator has been drained (bar_done)
                if (it_n > 0) {
                    mbar_wait(bar_o, ph ^ 1);
                    mbar_wait(bar_done, ph ^ 1);
                }
                mbar_arrive_expect_tx(bar_load, static_cast<uint32_t>(kTileQBytes + 2 * L * 128));
                tma_load_2d(&tmap_q, bar_load, s_q, h * kHd, row0 + qt * kQT, kCacheHintEvictFirst);
                for (int bx = 0; bx < full_kv_boxes; ++bx) {
                    tma_load_2d(&tmap_kv, bar_load, s_k + bx * kTileQBytes, W + h * kHd, row0 + bx * 128, kCacheHintEvictNormal);
                    tma_load_2d(&tmap_kv, bar_load, s_v + bx * kTileQBytes, 2 * W + h * kHd, row0 + bx * 128, kCacheHintEvictNormal);
                }
                if (kv_tail > 0) {
                    tma_load_2d(&tmap_kv_tail, bar_load, s_k + full_kv_boxes * kTileQBytes, W + h * kHd, row0 + full_kv_boxes * 128,
                                kCacheHintEvictNormal);
                    tma_load_2d(&tmap_kv_tail, bar_load, s_v + full_kv_boxes * kTileQBytes, 2 * W + h * kHd, row0 + full_kv_boxes * 128,
                                kCacheHintEvictNormal);
                }
                mbar_wait(bar_load, ph);
                tc_fence_after();
                // S = Q K^T
#pragma unroll
                for (int k = 0; k < kHd / 16; ++k) umma_f16(tmem_s, desc_q + 2 * k, desc_k + 2 * k, idesc_s1, k != 0);
                if (n2 > 0) {
#pragma unroll
                    for (int k = 0; k < kHd / 16; ++k)
                        umma_f16(tmem_s + 256, desc_q + 2 * k, desc_k + ((256u * 128u) >> 4) + 2 * k, idesc_s2, k != 0);
                }
                umma_commit(bar_s);
                // O = P V once the softmax threads have written P
                mbar_wait(bar_p, ph);
                tc_fence_after();
                const int ksteps = p.Lpad / 16;
                for (int j = 0; j < ksteps; ++j)
                    umma_f16_ts(tmem_o, tmem_p + 8 * j, desc_v + static_cast<uint64_t>(j) * ((16u * 128u) >> 4), idesc_o, j != 0);
                umma_commit(bar_o);
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax + output (thread = query row) =====================
        const int r = warp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
        const uint32_t so = smem_u32(s_o);
        const uint32_t row_off = static_cast<uint32_t>(r) * 128;
        const uint32_t rx = static_cast<uint32_t>(r & 7);
        const int nchunks = p.Lpad / 32;
        uint32_t it_n = 0;
        for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++it_n) {
            const uint32_t ph = it_n & 1;
            const int bh = it / p.q_tiles, qt = it - bh * p.q_tiles;
            const int b = bh / p.heads, h = bh - b * p.heads;
            const int qrow = qt * kQT + r;   // row inside the sequence (rows >= L are computed but never stored)
            mbar_wait(bar_s, ph);
            tc_fence_after();
            // pass 1: row maximum over the valid columns
            float mx = -INFINITY;
            for (int ch = 0; ch < nchunks; ++ch) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_s + lane_off + ch * 32, v);
                tmem_ld_wait();
                const int c0 = ch * 32;
                const bool need_mask = c0 + 32 > L || (CAUSAL && c0 + 32 > qrow + 1);
                if (need_mask) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c0 + j;
                        const float s = (col >= L || (CAUSAL && col > qrow)) ? -INFINITY : __uint_as_float(v[j]);
                        mx = fmaxf(mx, s);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                }
            }
            // column 0 is valid for every row (causal: col 0 <= row), so mx is finite unless the scores themselves are not
            const float nm = -mx * c;
            // pass 2: p = 2^(s*c - m*c), row sum, P (16-bit) over the S columns this thread has already consumed
            float sum = 0.f;
            for (int ch = 0; ch < nchunks; ++ch) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_s + lane_off + ch * 32, v);
                tmem_ld_wait();
                const int c0 = ch * 32;
                const bool need_mask = c0 + 32 > L || (CAUSAL && c0 + 32 > qrow + 1);
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float p0 = ex2_fast(fmaf(__uint_as_float(v[2 * j]), c, nm));
                    float p1 = ex2_fast(fmaf(__uint_as_float(v[2 * j + 1]), c, nm));
                    if (need_mask) {
                        const int col = c0 + 2 * j;
                        if (col >= L || (CAUSAL && col > qrow)) p0 = 0.f;
                        if (col + 1 >= L || (CAUSAL && col + 1 > qrow)) p1 = 0.f;
                    }
                    sum += p0 + p1;
                    pk[j] = H::pack(p0, p1);
                }
                tmem_st_32x16(tmem_p + lane_off + ch * 16, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_p);

            // O: TMEM -> registers -> * 1/sum -> swizzled staging -> TMA store
            mbar_wait(bar_o, ph);
            tc_fence_after();
            uint32_t o0[32], o1[32];
            tmem_ld_32x32(tmem_o + lane_off, o0);
            tmem_ld_32x32(tmem_o + lane_off + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar_done);
            const float inv = 1.0f / sum;
            if (tid == 0) tma_store_wait_read<0>();   // the previous item's store has released the staging tile
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const uint32_t* src = g < 4 ? &o0[g * 8] : &o1[(g - 4) * 8];
                uint4 w;
                w.x = H::pack(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
                w.y = H::pack(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
                w.z = H::pack(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
                w.w = H::pack(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
                const uint32_t addr = so + row_off + ((static_cast<uint32_t>(g) ^ rx) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
            }
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 0) {
                const bool tail = qt == p.q_tiles - 1 && p.tail_rows != kQT;
                tma_store_2d(tail ? &tmap_o_tail : &tmap_o, s_o, h * kHd, b * L + qt * kQT);
                tma_store_commit();
            }
        }
        if (tid == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    }
}

}  // namespace

// qkv [batch*L, 3*heads*64] -> out [batch*L, heads*64]; 16-bit dtypes, 64 < L <= 288.  Returns 1 when the shape is outside
// this kernel's range (the caller then uses the generic path), 0 on success, < 0 / CUDA code on error.
int attention_tc(int dtype, const void* qkv, void* out, int batch, int L, int heads, int causal, cudaStream_t stream) {
    if (!(dtype == 1 || dtype == 2) || L <= 64 || L > 288) return 1;
    const int W = heads * kHd;
    const int64_t rows = static_cast<int64_t>(batch) * L;
    AttnTcParams p;
    p.L = L;
    p.Lpad = (L + 31) / 32 * 32;
    p.heads = heads;
    p.q_tiles = (L + kQT - 1) / kQT;
    const int64_t items = static_cast<int64_t>(batch) * heads * p.q_tiles;
    B2C_CHECK_ARG(items <= 0x7fffffff, "attention: too many work items");
    p.items = static_cast<int>(items);
    p.tail_rows = L - (p.q_tiles - 1) * kQT;
    const int need = p.Lpad > p.Lpad / 2 + kHd ? p.Lpad : p.Lpad / 2 + kHd;
    p.tmem_cols = need <= 128 ? 128 : (need <= 256 ? 256 : 512);
    const int kv_tail = L % 128;

    const bool bf = dtype == 1;
    CUtensorMap tq, tkv, tkvt, to, tot;
    if (make_tmap_2d(&tq, bf, qkv, rows, 3 * W, 3 * W, kQT, kHd) != 0) return -1;
    tkv = tq;
    if (make_tmap_2d(&tkvt, bf, qkv, rows, 3 * W, 3 * W, kv_tail > 0 ? kv_tail : 128, kHd) != 0) return -1;
    if (make_tmap_2d(&to, bf, out, rows, W, W, kQT, kHd) != 0) return -1;
    if (make_tmap_2d(&tot, bf, out, rows, W, W, p.tail_rows, kHd) != 0) return -1;

    const int smem_bytes = 2 * kTileQBytes + 2 * p.Lpad * 128 + 64 + 1024;
    const int ctas_per_sm = p.tmem_cols <= 256 ? 2 : 1;
    const void* kerns[4] = {reinterpret_cast<const void*>(attention_tc_kernel<__nv_bfloat16, false>),
                            reinterpret_cast<const void*>(attention_tc_kernel<__nv_bfloat16, true>),
                            reinterpret_cast<const void*>(attention_tc_kernel<__half, false>),
                            reinterpret_cast<const void*>(attention_tc_kernel<__half, true>)};
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] {
        for (const void* k : kerns)
            if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(attention_tc smem)");
    B2C_CHECK_ARG(smem_bytes <= 120 * 1024, "attention_tc: shared memory budget exceeded");
    const int64_t max_ctas = static_cast<int64_t>(num_sms()) * ctas_per_sm;
    const int grid = static_cast<int>(items < max_ctas ? items : max_ctas);

    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(160);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t le;
    if (bf && !causal) le = cudaLaunchKernelEx(&cfg, attention_tc_kernel<__nv_bfloat16, false>, tq, tkv, tkvt, to, tot, p);
    else if (bf) le = cudaLaunchKernelEx(&cfg, attention_tc_kernel<__nv_bfloat16, true>, tq, tkv, tkvt, to, tot, p);
    else if (!causal) le = cudaLaunchKernelEx(&cfg, attention_tc_kernel<__half, false>, tq, tkv, tkvt, to, tot, p);
    else le = cudaLaunchKernelEx(&cfg, attention_tc_kernel<__half, true>, tq, tkv, tkvt, to, tot, p);
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(attention_tc_kernel)");
    B2C_LAUNCH_CHECK("attention_tc_kernel");
    return 0;
}

}  // namespace b200clip
