// Peer-memory exchange of the distributed ClipLoss (--local-loss --gather-with-grad, world_size > 1): replaces the two
// collectives of the reference step — torch.distributed.nn.all_gather of the features (deps/open_clip/src/open_clip/
// loss.py:49-50) and the reduce-scatter(SUM) of its autograd backward (torch/distributed/nn/functional.py:_AllGather) — by
// stores into the peers' HBM over NVLink / NVSwitch plus flag words, so a step is six small launches without a NCCL
// collective on its latency path:
//   forward   p2p_allgather_kernel   every rank converts its img | txt rows to fp32 and writes them into slot `rank` of EVERY
//                                    rank's gather buffer, raises its flag on each peer, then waits for all flags on itself
//             (packed forward kernels of cliploss.cu read the local gather buffer)
//   backward  the feature-gradient GEMM writes the [n, 2D] gradient block of rank j's rows straight into slot `rank` of rank
//             j's receive buffer (slot-addressed epilogue in cliploss.cu: the store IS the scatter)
//             p2p_reduce_finish_kernel   raises the "my blocks are written" flag on every peer, waits for all peers, sums
//                                        the `world` received blocks -> gradient of the local rows.
// The buffers are symmetric allocations mapped by the host (torch symmetric memory); this file only sees raw pointers.
// Flags carry a monotonically increasing epoch; a waiter accepts any value >= its epoch (a peer may already be one step on).
//
// Waits are bounded by WALL time (globaltimer), by default 600 s — the order of the NCCL watchdog, because rank skew of many
// seconds is routine in the reference's train loop (rank 0 evaluates and checkpoints while the others already run the next
// step, training/train.py:269).  A wait that expires does not trap (a trap is a sticky context error with no recovery): it
// raises the host-visible error word registered with b200clip_p2p_configure and lets the kernel retire; the host side
// (open_clip/peer.py) checks that word before every exchange and raises.
//
// Ring-slot protection across ranks: every rank owns one "busy" word per ring slot (= the epoch whose gathered rows in that
// slot are still needed by a pending backward, 0 when free).  A rank about to overwrite slot s of peer p for epoch e first
// waits until p's busy word is 0 or e, so a peer that runs ahead can never clobber features an un-backwarded forward still
// needs — it waits (and, if the slot is never released, reports the timeout) instead.
#include "common.cuh"
#include "internal.h"

#include <atomic>
#include <cstdlib>

namespace b200clip {

namespace {

std::atomic<unsigned long long> g_timeout_ns{0};       // 0 = not configured yet (env / default on first use)
std::atomic<uint32_t*> g_error_word{nullptr};

unsigned long long wait_timeout_ns() {
    unsigned long long v = g_timeout_ns.load(std::memory_order_relaxed);
    if (v == 0) {
        double s = 600.0;
        if (const char* e = getenv("B200CLIP_P2P_TIMEOUT_S")) {
            const double t = atof(e);
            if (t > 0.0) s = t;
        }
        v = static_cast<unsigned long long>(s * 1e9);
        g_timeout_ns.store(v, std::memory_order_relaxed);
    }
    return v;
}

struct WaitCfg {
    unsigned long long timeout_ns;
    uint32_t* err;    // device-accessible error word (pinned host memory), may be null
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void report_timeout(const WaitCfg& w, uint32_t code, uint32_t epoch) {
    printf("b200clip: peer wait timed out after %llu ms (code %u, block %d thread %d, epoch %u)\n", w.timeout_ns / 1000000ull, code,
           (int)blockIdx.x, (int)threadIdx.x, epoch);
    if (w.err != nullptr) {
        *reinterpret_cast<volatile uint32_t*>(w.err) = code;
        __threadfence_system();
    } else {
        __trap();      // nobody registered an error word: fail loudly rather than hand back undefined rows
    }
}
// bounded spin until *flag >= epoch (wrap-around safe).  A dead or badly skewed peer becomes a reported error, not a hung GPU.
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch, const WaitCfg& w) {
    if (static_cast<int32_t>(ld_acquire_sys(flag) - epoch) >= 0) return;
    const unsigned long long t0 = global_ns();
    while (static_cast<int32_t>(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(200);
        if (global_ns() - t0 > w.timeout_ns) {
            report_timeout(w, 1u, epoch);
            return;
        }
    }
}
// bounded spin until the owner of a ring slot has released it: busy word == 0 (free) or == epoch (the owner already runs this epoch)
__device__ __forceinline__ void wait_slot_free(const uint32_t* busy, uint32_t epoch, const WaitCfg& w) {
    uint32_t v = ld_acquire_sys(busy);
    if (v == 0u || v == epoch) return;
    const unsigned long long t0 = global_ns();
    while (true) {
        v = ld_acquire_sys(busy);
        if (v == 0u || v == epoch) return;
        __nanosleep(500);
        if (global_ns() - t0 > w.timeout_ns) {
            report_timeout(w, 2u, epoch);
            return;
        }
    }
}

template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = Half16<__nv_bfloat16>::unpack(u.x), b = Half16<__nv_bfloat16>::unpack(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 load4<__half>(const __half* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = Half16<__half>::unpack(u.x), b = Half16<__half>::unpack(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
}

// grid (chunks, world): block (c, p) writes chunk c of this rank's packed [n, 2D] fp32 rows into peer p's gather slot
template <typename T>
__global__ void __launch_bounds__(256)
p2p_allgather_kernel(const T* __restrict__ img, const T* __restrict__ txt, int n, int D, float* const* __restrict__ peer_dst,
                     uint32_t* const* __restrict__ peer_flag, const uint32_t* __restrict__ my_flags, uint32_t* counters, int world,
                     uint32_t epoch, uint32_t* const* __restrict__ peer_busy, uint32_t* my_busy, uint32_t hold, WaitCfg wc) {
    const int p = blockIdx.y;
    float* dst = peer_dst[p];
    // my own slot of this epoch: taken (a backward will read it) or free again as soon as this step's kernels are through
    if (my_busy != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) st_release_sys(my_busy, hold != 0u ? epoch : 0u);
    // peer p's slot must not hold rows an un-backwarded forward of p still needs
    if (peer_busy != nullptr) {
        if (threadIdx.x == 0) wait_slot_free(peer_busy[p], epoch, wc);
        __syncthreads();
    }
    const int quads_per_row = D / 2;   // float4 per packed row (2D floats)
    const int64_t total = static_cast<int64_t>(n) * quads_per_row;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / quads_per_row), q = static_cast<int>(i - static_cast<int64_t>(r) * quads_per_row);
        const int col = q * 4;
        const float4 v = col < D ? load4<T>(img + static_cast<int64_t>(r) * D + col) : load4<T>(txt + static_cast<int64_t>(r) * D + (col - D));
        *reinterpret_cast<float4*>(dst + static_cast<int64_t>(r) * 2 * D + col) = v;
    }
    // last block done for peer p publishes the slot
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counters + p, 1u);
        if (prev == gridDim.x - 1) {
            counters[p] = 0u;
            __threadfence_system();
            st_release_sys(peer_flag[p], epoch);
        }
    }
    // every rank's slot in MY buffer must have landed before the kernel (and with it the stream) moves on
    if (threadIdx.x < world) wait_flag(my_flags + threadIdx.x, epoch, wc);
    __syncthreads();
}

// recv [slots][elems] (slot q < world written by rank q's backward GEMM, slots >= world by the local one) -> out[elems] = sum
__global__ void __launch_bounds__(256)
p2p_reduce_finish_kernel(const float* __restrict__ recv, float* __restrict__ out, int64_t elems, uint32_t* const* __restrict__ peer_flag,
                         const uint32_t* __restrict__ my_flags, int world, int slots, uint32_t epoch, uint32_t* my_busy, int split_cols,
                         WaitCfg wc) {
    // the stores of the preceding kernel on this stream (the slot-addressed GEMM epilogue) are complete; publish them
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(peer_flag[threadIdx.x], epoch);
    }
    // that kernel was also the last reader of this step's gathered rows: the ring slot may be overwritten by the peers again
    if (my_busy != nullptr && blockIdx.x == 0 && threadIdx.x == 0) st_release_sys(my_busy, 0u);
    if (threadIdx.x < world) wait_flag(my_flags + threadIdx.x, epoch, wc);
    __syncthreads();
    const int64_t quads = elems / 4;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < quads; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float4 a = __ldcv(reinterpret_cast<const float4*>(recv) + i);
        for (int q = 1; q < slots; ++q) {
            const float4 b = __ldcv(reinterpret_cast<const float4*>(recv + static_cast<int64_t>(q) * elems) + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        if (split_cols > 0) {
            // rows of split_cols = 2D floats (img | txt gradient) -> two contiguous [rows, D] halves, so that the caller can hand
            // each half to autograd as a dense tensor (a column slice of [n, 2D] would be copied by AccumulateGrad)
            const int64_t e = i * 4;
            const int64_t row = e / split_cols;
            const int col = static_cast<int>(e - row * split_cols);
            const int half_cols = split_cols / 2;
            const int64_t dst = (col >= half_cols ? elems / 2 : 0) + row * half_cols + (col >= half_cols ? col - half_cols : col);
            *reinterpret_cast<float4*>(out + dst) = a;
        } else {
            reinterpret_cast<float4*>(out)[i] = a;
        }
    }
}

}  // namespace

int p2p_configure(double timeout_seconds, uint32_t* error_word) {
    B2C_CHECK_ARG(timeout_seconds > 0.0 && timeout_seconds < 1e7, "p2p_configure: timeout must be in (0, 1e7) seconds");
    g_timeout_ns.store(static_cast<unsigned long long>(timeout_seconds * 1e9), std::memory_order_relaxed);
    g_error_word.store(error_word, std::memory_order_relaxed);
    return 0;
}

int p2p_allgather(int dtype, const void* img, const void* txt, int n, int D, float* const* peer_dst, uint32_t* const* peer_flag,
                  const uint32_t* my_flags, uint32_t* counters, int world, uint32_t epoch, uint32_t* const* peer_busy, uint32_t* my_busy,
                  int hold, cudaStream_t stream) {
    const WaitCfg wc{wait_timeout_ns(), g_error_word.load(std::memory_order_relaxed)};
    const uint32_t hold_u = hold != 0 ? 1u : 0u;
    B2C_CHECK_ARG(img && txt && peer_dst && peer_flag && my_flags && counters, "p2p_allgather: null pointer");
    B2C_CHECK_ARG(n > 0 && D > 0 && D % 4 == 0 && world >= 1 && world <= 16, "p2p_allgather: bad shape n=%d D=%d world=%d", n, D, world);
    const int64_t quads = static_cast<int64_t>(n) * (D / 2);
    int chunks = static_cast<int>((quads + 2047) / 2048);            // 8 float4 per thread
    const int max_chunks = num_sms() / world > 0 ? num_sms() / world : 1;   // all blocks co-resident: the flag waits cannot starve a writer
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, world);
    if (dtype == 0)
        p2p_allgather_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(img), static_cast<const float*>(txt), n, D, peer_dst,
                                                              peer_flag, my_flags, counters, world, epoch, peer_busy, my_busy, hold_u, wc);
    else if (dtype == 1)
        p2p_allgather_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(img), static_cast<const __nv_bfloat16*>(txt),
                                                                      n, D, peer_dst, peer_flag, my_flags, counters, world, epoch, peer_busy, my_busy, hold_u, wc);
    else if (dtype == 2)
        p2p_allgather_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(img), static_cast<const __half*>(txt), n, D, peer_dst,
                                                               peer_flag, my_flags, counters, world, epoch, peer_busy, my_busy, hold_u, wc);
    else
        B2C_CHECK_ARG(false, "p2p_allgather: bad dtype %d", dtype);
    B2C_LAUNCH_CHECK("p2p_allgather_kernel");
    return 0;
}

int p2p_reduce_finish(const float* recv, float* out, int64_t elems, uint32_t* const* peer_flag, const uint32_t* my_flags, int world,
                      int slots, uint32_t epoch, uint32_t* my_busy, int split_cols, cudaStream_t stream) {
    const WaitCfg wc{wait_timeout_ns(), g_error_word.load(std::memory_order_relaxed)};
    B2C_CHECK_ARG(recv && out && peer_flag && my_flags, "p2p_reduce_finish: null pointer");
    B2C_CHECK_ARG(elems > 0 && elems % 4 == 0 && world >= 1 && world <= 16 && slots >= world,
                  "p2p_reduce_finish: bad shape elems=%lld world=%d slots=%d", static_cast<long long>(elems), world, slots);
    B2C_CHECK_ARG(split_cols == 0 || (split_cols > 0 && split_cols % 8 == 0 && elems % split_cols == 0),
                  "p2p_reduce_finish: split_cols=%d must be 0 or a multiple of 8 that divides elems", split_cols);
    int blocks = static_cast<int>((elems / 4 + 1023) / 1024);
    if (blocks > num_sms()) blocks = num_sms();
    if (blocks < 1) blocks = 1;
    p2p_reduce_finish_kernel<<<blocks, 256, 0, stream>>>(recv, out, elems, peer_flag, my_flags, world, slots, epoch, my_busy, split_cols, wc);
    B2C_LAUNCH_CHECK("p2p_reduce_finish_kernel");
    return 0;
}

}  // namespace b200clip
