// ClipLoss forward + backward for the --local-loss --gather-with-grad configuration
// (deps/open_clip/src/open_clip/loss.py:102-131, labels :89-100):
//   logits_per_image = s * img_loc @ all_txt^T          [n, N]
//   logits_per_text  = s * txt_loc @ all_img^T          [n, N]
//   labels_i = i + n * rank;   loss = (CE(logits_per_image) + CE(logits_per_text)) / 2
// world_size == 1 is the same code with N == n and all_* == *_loc.
//
// The whole step is a few GFLOP on [256, 2048]-sized operands: it is bound by launch latency, not by any pipe.
// So the structure is FOUR launches for forward + backward (the reference's eager path issues ~25):
//   1. both logit blocks in one batched fp32 GEMM launch (64 x 64 tiles -> enough CTAs even at N = 256),
//   2. row-wise log-sum-exp / cross-entropy, the loss reduction folded in through a last-CTA-done counter
//      (deterministic: the last CTA sums the per-row terms in a fixed order),
//   3. d(logits) in place (scaled by the upstream gradient read from device memory) + d(logit_scale), same fold,
//   4. all four feature-gradient GEMMs in one batched launch.
// fp32 FFMA with fp32 accumulation throughout: the loss / gradient parity gates are fp32 gates.
#include "common.cuh"
#include "internal.h"

namespace b200clip {

namespace {

constexpr int kT = 64;    // output tile edge
constexpr int kTK = 16;   // k step

struct GemmProb {
    const float* A;   // ta == 0: [M, K] (K contiguous);  ta == 1: stored [K, M]
    const float* B;   // tb == 0: [N, K] (K contiguous);  tb == 1: stored [K, N]
    float* C;         // [M, N]
    int64_t lda, ldb, ldc;
    int M, N, K, ta, tb;
    int accumulate;   // C += A B instead of C = A B
    // optional second contraction segment, summed into the same output tile before it is stored: C = A B + A2 B2 with A2 stored
    // [K2, M] and B2 stored [K2, N] (both "transposed" forms).  The single-device backward uses it to write the TOTAL gradient of
    // a feature matrix that is both the row and the column operand of the logit blocks (K2 == 0: none).
    const float* A2;
    const float* B2;
    int64_t lda2, ldb2;
    int K2;
    int tiles_n, tile_begin;
    // slot-addressed output (peer-memory scatter of the distributed backward, p2p.cu): row r of C lives at
    // slots[(slot_row0 + r) / slot_rows] + ((slot_row0 + r) % slot_rows) * ldc + slot_col0; C itself is unused then
    float* const* slots;
    int slot_rows, slot_row0, slot_col0;
};
constexpr int kMaxProbs = 6;
struct GemmBatch {
    GemmProb prob[kMaxProbs];
    int count;
    unsigned int* zero4;   // optional: four words cleared by the first CTA (the last-CTA-done counter of the kernel that follows)
};

// One k-step of an operand tile: [kTK x 64] in shared memory, k-major.  The global read and the shared-memory write are
// separate so that the next k-step's global loads are in flight while this one is multiplied (register double buffering).
// operand stored with the contraction index contiguous ([rows, K]) ...
__device__ __forceinline__ float4 load_kc(const float* base, int64_t ld, int r0, int rows, int k0, int K, int tid) {
    const int r = tid >> 2;
    const int lk = (tid & 3) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < rows && k0 + lk < K) {
        const float* src = base + static_cast<int64_t>(r0 + r) * ld + k0 + lk;
        if (k0 + lk + 3 < K && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            a = *reinterpret_cast<const float4*>(src);
        } else {   // ragged K or an unaligned row: element by element
            a.x = src[0];
            if (k0 + lk + 1 < K) a.y = src[1];
            if (k0 + lk + 2 < K) a.z = src[2];
            if (k0 + lk + 3 < K) a.w = src[3];
        }
    }
    return a;
}
__device__ __forceinline__ void store_kc(float (*S)[kT + 4], float4 a, int tid) {
    const int r = tid >> 2;
    const int lk = (tid & 3) * 4;
    S[lk + 0][r] = a.x; S[lk + 1][r] = a.y; S[lk + 2][r] = a.z; S[lk + 3][r] = a.w;
}
// ... or with the output index contiguous ([K, rows])
__device__ __forceinline__ float4 load_mc(const float* base, int64_t ld, int r0, int rows, int k0, int K, int tid) {
    const int k = tid >> 4;
    const int r = (tid & 15) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + k < K && r0 + r < rows) {
        const float* src = base + static_cast<int64_t>(k0 + k) * ld + r0 + r;
        if (r0 + r + 3 < rows && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            a = *reinterpret_cast<const float4*>(src);
        } else {   // ragged row count or an unaligned row: element by element
            a.x = src[0];
            if (r0 + r + 1 < rows) a.y = src[1];
            if (r0 + r + 2 < rows) a.z = src[2];
            if (r0 + r + 3 < rows) a.w = src[3];
        }
    }
    return a;
}
__device__ __forceinline__ void store_mc(float (*S)[kT + 4], float4 a, int tid) {
    const int k = tid >> 4;
    const int r = (tid & 15) * 4;
    *reinterpret_cast<float4*>(&S[k][r]) = a;
}

// acc += op(A)[m0.., :K] op(B)[n0.., :K]^T for one 64 x 64 output tile (thread (ty, tx) owns a 4 x 4 block)
template <bool TA, bool TB>
__device__ __forceinline__ void tile_accumulate(const float* A, int64_t lda, int M, const float* B, int64_t ldb, int N, int K, int m0, int n0,
                                                float (*As)[kT + 4], float (*Bs)[kT + 4], uint64_t (&acc)[4][2]) {
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    auto load_a = [&](int k0) { return TA ? load_mc(A, lda, m0, M, k0, K, tid) : load_kc(A, lda, m0, M, k0, K, tid); };
    auto load_b = [&](int k0) { return TB ? load_mc(B, ldb, n0, N, k0, K, tid) : load_kc(B, ldb, n0, N, k0, K, tid); };
    float4 ra = load_a(0), rb = load_b(0);
    for (int k0 = 0; k0 < K; k0 += kTK) {
        if constexpr (TA) store_mc(As, ra, tid);
        else store_kc(As, ra, tid);
        if constexpr (TB) store_mc(Bs, rb, tid);
        else store_kc(Bs, rb, tid);
        __syncthreads();
        if (k0 + kTK < K) {   // next k-step's operands: in flight under the FMAs below
            ra = load_a(k0 + kTK);
            rb = load_b(k0 + kTK);
        }
#pragma unroll
        for (int k = 0; k < kTK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            // packed fp32x2 FMAs (two columns per instruction; the same round-to-nearest fused arithmetic as fmaf): the inner
            // loop is bound by FMA issue slots, so this is worth 1.6-1.8x on the whole kernel
            const uint64_t b01 = pack_f2(b.x, b.y), b23 = pack_f2(b.z, b.w);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint64_t aa = pack_f2(av[i], av[i]);
                acc[i][0] = fma_f2(aa, b01, acc[i][0]);
                acc[i][1] = fma_f2(aa, b23, acc[i][1]);
            }
        }
        __syncthreads();
    }
}

template <bool TA, bool TB>
__device__ __forceinline__ void tile_gemm(const GemmProb& p, int m0, int n0, float (*As)[kT + 4], float (*Bs)[kT + 4]) {
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    uint64_t acc2[4][2];   // 4 x 4 fp32 accumulators as packed column pairs
#pragma unroll
    for (int i = 0; i < 4; ++i) acc2[i][0] = acc2[i][1] = pack_f2(0.f, 0.f);
    tile_accumulate<TA, TB>(p.A, p.lda, p.M, p.B, p.ldb, p.N, p.K, m0, n0, As, Bs, acc2);
    if (p.K2 > 0) tile_accumulate<true, true>(p.A2, p.lda2, p.M, p.B2, p.ldb2, p.N, p.K2, m0, n0, As, Bs, acc2);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unpack_f2(acc2[i][0], acc[i][0], acc[i][1]);
        unpack_f2(acc2[i][1], acc[i][2], acc[i][3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        const int col = n0 + tx * 4;
        if (row < p.M && col < p.N) {
            float* dstf;
            if (p.slots != nullptr) {
                const int gr = p.slot_row0 + row;
                const int sl = gr / p.slot_rows;
                dstf = p.slots[sl] + static_cast<int64_t>(gr - sl * p.slot_rows) * p.ldc + p.slot_col0 + col;
            } else {
                dstf = p.C + static_cast<int64_t>(row) * p.ldc + col;
            }
            if (col + 3 < p.N && (reinterpret_cast<uintptr_t>(dstf) & 15) == 0) {
                float4* dst = reinterpret_cast<float4*>(dstf);
                float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                if (p.accumulate) {
                    const float4 c = *dst;
                    o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
                }
                *dst = o;
            } else {   // ragged N or an unaligned row
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < p.N) dstf[j] = p.accumulate ? dstf[j] + acc[i][j] : acc[i][j];
            }
        }
    }
}

__global__ void __launch_bounds__(256) batched_gemm_kernel(const GemmBatch batch) {
    __shared__ __align__(16) float As[kTK][kT + 4];
    __shared__ __align__(16) float Bs[kTK][kT + 4];
    if (batch.zero4 != nullptr && blockIdx.x == 0 && threadIdx.x < 4) batch.zero4[threadIdx.x] = 0u;
    int pi = 0;
#pragma unroll
    for (int i = 1; i < kMaxProbs; ++i)
        if (i < batch.count && static_cast<int>(blockIdx.x) >= batch.prob[i].tile_begin) pi = i;
    const GemmProb& p = batch.prob[pi];
    const int t = blockIdx.x - p.tile_begin;
    const int m0 = (t / p.tiles_n) * kT, n0 = (t % p.tiles_n) * kT;
    if (!p.ta && !p.tb) tile_gemm<false, false>(p, m0, n0, As, Bs);
    else if (!p.ta && p.tb) tile_gemm<false, true>(p, m0, n0, As, Bs);
    else if (p.ta && !p.tb) tile_gemm<true, false>(p, m0, n0, As, Bs);
    else tile_gemm<true, true>(p, m0, n0, As, Bs);
}

int launch_batch(GemmBatch& b, cudaStream_t stream) {
    int tiles = 0;
    for (int i = 0; i < b.count; ++i) {
        GemmProb& p = b.prob[i];
        p.tiles_n = (p.N + kT - 1) / kT;
        p.tile_begin = tiles;
        tiles += p.tiles_n * ((p.M + kT - 1) / kT);
    }
    batched_gemm_kernel<<<tiles, 256, 0, stream>>>(b);
    B2C_LAUNCH_CHECK("cliploss batched_gemm_kernel");
    return 0;
}

GemmProb make_prob(const float* A, int64_t lda, bool ta, const float* B, int64_t ldb, bool tb, float* C, int64_t ldc, int M, int N, int K,
                   bool accumulate = false) {
    GemmProb p;
    p.A = A; p.B = B; p.C = C;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K;
    p.ta = ta ? 1 : 0; p.tb = tb ? 1 : 0;
    p.accumulate = accumulate ? 1 : 0;
    p.A2 = nullptr; p.B2 = nullptr; p.lda2 = 0; p.ldb2 = 0; p.K2 = 0;
    p.tiles_n = 0; p.tile_begin = 0;
    p.slots = nullptr; p.slot_rows = 1; p.slot_row0 = 0; p.slot_col0 = 0;
    return p;
}
// + A2 B2 with A2 stored [K2, M], B2 stored [K2, N]
GemmProb with_second_segment(GemmProb p, const float* A2, int64_t lda2, const float* B2, int64_t ldb2, int K2) {
    p.A2 = A2; p.B2 = B2; p.lda2 = lda2; p.ldb2 = ldb2; p.K2 = K2;
    return p;
}
GemmProb to_slots(GemmProb p, float* const* slots, int slot_rows, int row0, int col0) {
    p.slots = slots; p.slot_rows = slot_rows; p.slot_row0 = row0; p.slot_col0 = col0;
    return p;
}

}  // namespace

// General fp32 GEMM with optional operand transposes, the kernel of the loss step reused by the fp32 (parity-mode) tower
// backward: C[M,N] (+)= op(A) op(B)^T-free form  C = A' B' with A' = ta ? A^T : A (A stored [K,M] when ta) and
// B' = tb ? B : B^T (B stored [K,N] when tb, [N,K] otherwise).  Any shape (16-byte vector accesses where a row allows them).
int gemm_f32_general(const float* A, int64_t lda, bool ta, const float* B, int64_t ldb, bool tb, float* C, int64_t ldc, int M, int N, int K,
                     bool accumulate, cudaStream_t stream) {
    B2C_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "gemm_f32_general: bad arguments");
    GemmBatch b;
    b.zero4 = nullptr;
    b.count = 1;
    b.prob[0] = make_prob(A, lda, ta, B, ldb, tb, C, ldc, M, N, K, accumulate);
    return launch_batch(b, stream);
}

namespace {

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
}

// True in exactly one CTA of the grid: the one that finishes last.  `counter` must be zero on entry and is reset to zero.
__device__ __forceinline__ bool last_block_done(unsigned int* counter) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counter, 1u);
        is_last = prev == gridDim.x - 1;
        if (is_last) *counter = 0u;
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

// workspace layout (floats): Li [n*N] | Lt [n*N] | row_loss [2n] | row_lse [2n] | row_ds [2n] | counter [4]
struct Ws {
    float *logits, *row_loss, *row_lse, *row_ds;
    unsigned int* counter;
};
Ws carve(float* w, int n, int N) {
    Ws s;
    s.logits = w;
    s.row_loss = w + 2 * static_cast<int64_t>(n) * N;
    s.row_lse = s.row_loss + 2 * n;
    s.row_ds = s.row_lse + 2 * n;
    s.counter = reinterpret_cast<unsigned int*>(s.row_ds + 2 * n);
    return s;
}

// rows [0,n): image->text logits; rows [n,2n): text->image logits.  `x` holds raw dot products (kept for the backward).
__global__ void __launch_bounds__(256)
ce_forward_kernel(const float* __restrict__ x, const float* __restrict__ logit_scale, int n, int N, int rank,
                  float* __restrict__ row_loss, float* __restrict__ row_lse, unsigned int* counter, float* __restrict__ loss) {
    __shared__ float red[8];
    const int r = blockIdx.x;
    const float* xr = x + static_cast<int64_t>(r) * N;
    const float s = *logit_scale;
    const int label = (r % n) + n * rank;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < N; j += blockDim.x) mx = fmaxf(mx, s * xr[j]);
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) sum += expf(s * xr[j] - mx);
    sum = block_reduce(sum, red, false);
    const float lse = mx + logf(sum);
    if (threadIdx.x == 0) {
        row_lse[r] = lse;
        row_loss[r] = lse - s * xr[label];
    }
    if (last_block_done(counter)) {
        float a = 0.f;
        for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) a += __ldcg(row_loss + i);
        a = block_reduce(a, red, false);
        if (threadIdx.x == 0) *loss = a / (2.f * static_cast<float>(n));
    }
}

// x[r, j] <- s * dL/dlogit[r, j] in place;  d_scale = sum_rj dL/dlogit * raw (folded in like the loss)
__global__ void __launch_bounds__(256)
ce_backward_kernel(float* __restrict__ x, const float* __restrict__ logit_scale, const float* __restrict__ grad_out, int n, int N,
                   int rank, const float* __restrict__ row_lse, float* __restrict__ row_ds, unsigned int* counter,
                   float* __restrict__ d_scale) {
    __shared__ float red[8];
    const int r = blockIdx.x;
    float* xr = x + static_cast<int64_t>(r) * N;
    const float s = *logit_scale;
    const int label = (r % n) + n * rank;
    const float lse = row_lse[r];
    const float g = (grad_out != nullptr ? *grad_out : 1.f) / (2.f * static_cast<float>(n));
    float ds = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float raw = xr[j];
        const float p = expf(s * raw - lse);
        const float dl = g * (p - (j == label ? 1.f : 0.f));
        ds += dl * raw;
        xr[j] = s * dl;
    }
    ds = block_reduce(ds, red, false);
    if (threadIdx.x == 0) row_ds[r] = ds;
    if (last_block_done(counter)) {
        float b = 0.f;
        for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) b += __ldcg(row_ds + i);
        b = block_reduce(b, red, false);
        if (threadIdx.x == 0 && d_scale != nullptr) *d_scale = b;
    }
}

int check_shapes(int rank, int n, int N, int D) {
    B2C_CHECK_ARG(n > 0 && N >= n && D > 0 && N % n == 0, "cliploss: bad shape n=%d N=%d D=%d", n, N, D);
    B2C_CHECK_ARG(rank >= 0 && (rank + 1) * n <= N, "cliploss: rank %d out of range for n=%d N=%d", rank, n, N);
    return 0;   // any n, D: the GEMM tiles fall back to scalar accesses on ragged / unaligned rows
}

// Operand views with explicit row pitches: the contiguous API passes pitch D everywhere, the packed (img | txt) API of the
// distributed path passes pitch 2D and column offsets.
struct LossOperands {
    const float *img_loc, *txt_loc, *all_img, *all_txt;
    int64_t ld_loc, ld_all;
};

int forward_impl(const LossOperands& o, const float* logit_scale, int rank, int n, int N, int D, float* loss, float* workspace,
                 cudaStream_t stream) {
    int rc;
    if ((rc = check_shapes(rank, n, N, D)) != 0) return rc;
    B2C_CHECK_ARG(o.img_loc && o.txt_loc && o.all_img && o.all_txt && logit_scale && loss && workspace, "cliploss: null pointer");
    const Ws ws = carve(workspace, n, N);
    GemmBatch b;
    b.zero4 = ws.counter;   // last-CTA-done counter of the two cross-entropy kernels (each leaves it at zero again): cleared by the
                            // GEMM launch that precedes them instead of a separate memset
    b.count = 2;
    b.prob[0] = make_prob(o.img_loc, o.ld_loc, false, o.all_txt, o.ld_all, false, ws.logits, N, n, N, D);
    b.prob[1] = make_prob(o.txt_loc, o.ld_loc, false, o.all_img, o.ld_all, false, ws.logits + static_cast<int64_t>(n) * N, N, n, N, D);
    if ((rc = launch_batch(b, stream)) != 0) return rc;
    ce_forward_kernel<<<2 * n, 256, 0, stream>>>(ws.logits, logit_scale, n, N, rank, ws.row_loss, ws.row_lse, ws.counter, loss);
    B2C_LAUNCH_CHECK("ce_forward_kernel");
    return 0;
}

// d_*_loc == nullptr with `fold_local`: the local-row gradients are ADDED into rows [rank*n, rank*n + n) of d_all_* (second
// launch), which is what a following reduce-scatter needs
int backward_impl(const LossOperands& o, const float* logit_scale, int rank, int n, int N, int D, const float* grad_out,
                  float* d_img_loc, float* d_txt_loc, int64_t ld_dloc, float* d_all_img, float* d_all_txt, int64_t ld_dall,
                  bool fold_local, float* d_scale, float* workspace, cudaStream_t stream, float* const* d_slots = nullptr) {
    int rc;
    if ((rc = check_shapes(rank, n, N, D)) != 0) return rc;
    B2C_CHECK_ARG(o.img_loc && o.txt_loc && o.all_img && o.all_txt && logit_scale && workspace, "cliploss: null pointer");
    const Ws ws = carve(workspace, n, N);
    ce_backward_kernel<<<2 * n, 256, 0, stream>>>(ws.logits, logit_scale, grad_out, n, N, rank, ws.row_lse, ws.row_ds, ws.counter, d_scale);
    B2C_LAUNCH_CHECK("ce_backward_kernel");
    // the workspace now holds s * dL/dlogits
    float* Li = ws.logits;
    float* Lt = ws.logits + static_cast<int64_t>(n) * N;
    GemmBatch b;
    b.zero4 = nullptr;
    b.count = 0;
    if (!fold_local) {
        if (d_img_loc) b.prob[b.count++] = make_prob(Li, N, false, o.all_txt, o.ld_all, true, d_img_loc, ld_dloc, n, D, N);
        if (d_txt_loc) b.prob[b.count++] = make_prob(Lt, N, false, o.all_img, o.ld_all, true, d_txt_loc, ld_dloc, n, D, N);
    }
    if (d_slots != nullptr) {
        // Peer-memory scatter, ONE launch.  The local-row terms (contraction over all N gathered rows: few tiles, long K) come
        // first in the grid and are split in two K halves, each written to its own local slot (d_slots[world], [world + 1]);
        // the gathered-row terms (many tiles, K = n) follow and store rank j's [n, 2D] block to d_slots[j] in j's memory.
        // b200clip_p2p_reduce_finish sums all world + 2 slots, so nothing is accumulated in place and nothing is ordered.
        B2C_CHECK_ARG(fold_local, "cliploss: the slot-addressed backward is the packed (fold_local) form");
        const int world = N / n;
        const int kh = (N / 2 + 3) / 4 * 4;   // K split point (multiple of 4: float4 staging)
        for (int h = 0; h < 2; ++h) {
            const int k0 = h * kh, kn = h == 0 ? kh : N - kh;
            if (kn <= 0) continue;
            b.prob[b.count++] = to_slots(make_prob(Li + k0, N, false, o.all_txt + static_cast<int64_t>(k0) * o.ld_all, o.ld_all, true, nullptr,
                                                   ld_dall, n, D, kn), d_slots, n, (world + h) * n, 0);
            b.prob[b.count++] = to_slots(make_prob(Lt + k0, N, false, o.all_img + static_cast<int64_t>(k0) * o.ld_all, o.ld_all, true, nullptr,
                                                   ld_dall, n, D, kn), d_slots, n, (world + h) * n, D);
        }
        b.prob[b.count++] = to_slots(make_prob(Li, N, true, o.img_loc, o.ld_loc, true, nullptr, ld_dall, N, D, n), d_slots, n, 0, D);
        b.prob[b.count++] = to_slots(make_prob(Lt, N, true, o.txt_loc, o.ld_loc, true, nullptr, ld_dall, N, D, n), d_slots, n, 0, 0);
        return launch_batch(b, stream);
    }
    if (d_all_txt) b.prob[b.count++] = make_prob(Li, N, true, o.img_loc, o.ld_loc, true, d_all_txt, ld_dall, N, D, n);
    if (d_all_img) b.prob[b.count++] = make_prob(Lt, N, true, o.txt_loc, o.ld_loc, true, d_all_img, ld_dall, N, D, n);
    if (b.count > 0 && (rc = launch_batch(b, stream)) != 0) return rc;
    if (fold_local) {
        B2C_CHECK_ARG(d_all_img && d_all_txt, "cliploss: folding the local gradients needs d_all_img and d_all_txt");
        GemmBatch f;
        f.zero4 = nullptr;
        f.count = 2;
        f.prob[0] = make_prob(Li, N, false, o.all_txt, o.ld_all, true, d_all_img + static_cast<int64_t>(rank) * n * ld_dall, ld_dall, n, D, N, true);
        f.prob[1] = make_prob(Lt, N, false, o.all_img, o.ld_all, true, d_all_txt + static_cast<int64_t>(rank) * n * ld_dall, ld_dall, n, D, N, true);
        if ((rc = launch_batch(f, stream)) != 0) return rc;
    }
    return 0;
}

}  // namespace

int cliploss_forward(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt, const float* logit_scale,
                     int rank, int n, int N, int D, float* loss, float* workspace, cudaStream_t stream) {
    const LossOperands o{img_loc, txt_loc, all_img, all_txt, D, D};
    return forward_impl(o, logit_scale, rank, n, N, D, loss, workspace, stream);
}

int cliploss_backward(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt, const float* logit_scale,
                      int rank, int n, int N, int D, const float* grad_out, float* d_img_loc, float* d_txt_loc, float* d_all_img,
                      float* d_all_txt, float* d_scale, float* workspace, cudaStream_t stream) {
    const LossOperands o{img_loc, txt_loc, all_img, all_txt, D, D};
    return backward_impl(o, logit_scale, rank, n, N, D, grad_out, d_img_loc, d_txt_loc, D, d_all_img, d_all_txt, D, false, d_scale,
                         workspace, stream);
}

// Packed layout of the distributed path: `gathered` [N, 2D] holds img | txt of every rank (the ONE all-gather payload), the
// local rows are rows [rank*n, rank*n + n) of it.  backward writes d_gathered [N, 2D] = gradient w.r.t. every gathered row
// INCLUDING the local-row terms, i.e. exactly the input of the reduce-scatter that finishes gather_with_grad.
int cliploss_packed_forward(const float* gathered, const float* logit_scale, int rank, int n, int N, int D, float* loss,
                            float* workspace, cudaStream_t stream) {
    B2C_CHECK_ARG(gathered != nullptr, "cliploss: null pointer");
    const float* loc = gathered + static_cast<int64_t>(rank) * n * 2 * D;
    const LossOperands o{loc, loc + D, gathered, gathered + D, 2 * static_cast<int64_t>(D), 2 * static_cast<int64_t>(D)};
    return forward_impl(o, logit_scale, rank, n, N, D, loss, workspace, stream);
}

int cliploss_packed_backward(const float* gathered, const float* logit_scale, int rank, int n, int N, int D, const float* grad_out,
                             float* d_gathered, float* d_scale, float* workspace, cudaStream_t stream) {
    B2C_CHECK_ARG(gathered != nullptr && d_gathered != nullptr, "cliploss: null pointer");
    const float* loc = gathered + static_cast<int64_t>(rank) * n * 2 * D;
    const LossOperands o{loc, loc + D, gathered, gathered + D, 2 * static_cast<int64_t>(D), 2 * static_cast<int64_t>(D)};
    return backward_impl(o, logit_scale, rank, n, N, D, grad_out, nullptr, nullptr, 0, d_gathered, d_gathered + D,
                         2 * static_cast<int64_t>(D), true, d_scale, workspace, stream);
}

// Same, with the reduce-scatter's scatter half folded into the GEMM epilogue: d_slots[j], j < world = N/n, is where the [n, 2D]
// gradient block of rank j's rows goes — slot `rank` of rank j's receive buffer in peer memory; d_slots[world] and
// d_slots[world + 1] are two LOCAL [n, 2D] slots that receive the two K halves of the local-row terms.
int cliploss_packed_backward_p2p(const float* gathered, const float* logit_scale, int rank, int n, int N, int D, const float* grad_out,
                                 float* const* d_slots, float* d_scale, float* workspace, cudaStream_t stream) {
    B2C_CHECK_ARG(gathered != nullptr && d_slots != nullptr, "cliploss: null pointer");
    const float* loc = gathered + static_cast<int64_t>(rank) * n * 2 * D;
    const LossOperands o{loc, loc + D, gathered, gathered + D, 2 * static_cast<int64_t>(D), 2 * static_cast<int64_t>(D)};
    return backward_impl(o, logit_scale, rank, n, N, D, grad_out, nullptr, nullptr, 0, nullptr, nullptr, 2 * static_cast<int64_t>(D), true,
                         d_scale, workspace, stream, d_slots);
}

// Single-device step (world_size == 1: the feature matrices are both the row and the column operands of the two logit blocks,
// loss.py:102-131 with all_* == *_loc): backward of `cliploss_forward(img, txt, img, txt, ..., rank 0, n, n, D, ...)` that writes
// the TOTAL gradients  d_img = Li txt + Lt^T txt,  d_txt = Lt img + Li^T img  (Li, Lt = s dL/dlogits of the two blocks) with one
// two-segment GEMM launch: no separate row / column gradients for autograd to add up afterwards.
int cliploss_single_backward(const float* img, const float* txt, const float* logit_scale, int n, int D, const float* grad_out,
                             float* d_img, float* d_txt, float* d_scale, float* workspace, cudaStream_t stream) {
    int rc;
    if ((rc = check_shapes(0, n, n, D)) != 0) return rc;
    B2C_CHECK_ARG(img && txt && logit_scale && workspace, "cliploss: null pointer");
    const Ws ws = carve(workspace, n, n);
    ce_backward_kernel<<<2 * n, 256, 0, stream>>>(ws.logits, logit_scale, grad_out, n, n, 0, ws.row_lse, ws.row_ds, ws.counter, d_scale);
    B2C_LAUNCH_CHECK("ce_backward_kernel");
    const float* Li = ws.logits;
    const float* Lt = ws.logits + static_cast<int64_t>(n) * n;
    GemmBatch b;
    b.zero4 = nullptr;
    b.count = 0;
    if (d_img) b.prob[b.count++] = with_second_segment(make_prob(Li, n, false, txt, D, true, d_img, D, n, D, n), Lt, n, txt, D, n);
    if (d_txt) b.prob[b.count++] = with_second_segment(make_prob(Lt, n, false, img, D, true, d_txt, D, n, D, n), Li, n, img, D, n);
    return b.count > 0 ? launch_batch(b, stream) : 0;
}

// fused forward + backward (one call; used when the upstream gradient is already known or 1)
int cliploss(const float* img_loc, const float* txt_loc, const float* all_img, const float* all_txt, const float* logit_scale,
             int rank, int n, int N, int D, float* loss, const float* grad_out, float* d_img_loc, float* d_txt_loc,
             float* d_all_img, float* d_all_txt, float* d_scale, float* workspace, cudaStream_t stream) {
    int rc;
    B2C_CHECK_ARG(workspace != nullptr, "cliploss: null workspace");
    if ((rc = check_shapes(rank, n, N, D)) != 0) return rc;
    if ((rc = cliploss_forward(img_loc, txt_loc, all_img, all_txt, logit_scale, rank, n, N, D, loss, workspace, stream)) != 0) return rc;
    const bool want_grad = d_img_loc || d_txt_loc || d_all_img || d_all_txt || d_scale;
    if (!want_grad) return 0;
    return cliploss_backward(img_loc, txt_loc, all_img, all_txt, logit_scale, rank, n, N, D, grad_out, d_img_loc, d_txt_loc, d_all_img,
                             d_all_txt, d_scale, workspace, stream);
}

}  // namespace b200clip
