// Shared device/host helpers for the b200clip sm_100a kernels.
//
// Everything here is hand-written inline PTX for Blackwell (sm_100a): mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the small math helpers
// the epilogues share.  No CUTLASS / CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b200clip {

// ---------------------------------------------------------------------------------------
// Error plumbing (no exceptions cross the C ABI; see include/b200clip.h)
// ---------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n);

#define B2C_CHECK_ARG(cond, ...)                    \
    do {                                            \
        if (!(cond)) {                              \
            ::b200clip::set_last_error(__VA_ARGS__);\
            return -1;                              \
        }                                           \
    } while (0)

#define B2C_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return ::b200clip::cuda_fail(e__, #call);  \
    } while (0)

#define B2C_LAUNCH_CHECK(name)                                             \
    do {                                                                   \
        cudaError_t e__ = cudaGetLastError();                              \
        if (e__ != cudaSuccess) return ::b200clip::cuda_fail(e__, name);   \
        ::b200clip::count_launch(1);                                       \
    } while (0)

int num_sms();
// false when B200CLIP_NO_PDL=1: kernels are then launched without the programmatic-dependent-launch attribute
bool pdl_enabled();

// ---------------------------------------------------------------------------------------
// 16-bit storage types: bf16 (F=1) and fp16 (F=0) share every kernel through this trait.
// ---------------------------------------------------------------------------------------
template <typename T> struct Half16;
template <> struct Half16<__nv_bfloat16> {
    static constexpr uint32_t kUmmaFormat = 1;  // tcgen05 instruction descriptor a/b format: BF16
    using T2 = __nv_bfloat162;
    __device__ __forceinline__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ __forceinline__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
    __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
        __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&p);
    }
    __device__ __forceinline__ static float2 unpack(uint32_t u) {
        __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162*>(&u);
        return __bfloat1622float2(p);
    }
};
template <> struct Half16<__half> {
    static constexpr uint32_t kUmmaFormat = 0;  // F16
    using T2 = __half2;
    __device__ __forceinline__ static float to_f(__half v) { return __half2float(v); }
    __device__ __forceinline__ static __half from_f(float v) { return __float2half_rn(v); }
    __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
        __half2 p = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&p);
    }
    __device__ __forceinline__ static float2 unpack(uint32_t u) {
        __half2 p = *reinterpret_cast<__half2*>(&u);
        return __half22float2(p);
    }
};

// round-trip through the 16-bit storage type (mirrors the reference's per-op rounding points)
template <typename T> __device__ __forceinline__ float round16(float v) {
    return Half16<T>::to_f(Half16<T>::from_f(v));
}

// ---------------------------------------------------------------------------------------
// Activations (open_clip transformer.py:33-36 QuickGELU; nn.GELU default = exact erf GELU)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }
// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7 + fp32 round-off): 2 MUFU + ~10 FP32 ops instead of erff's ~25
// branching instructions.  Used only where the result is rounded to bf16/fp16 anyway (the GEMM epilogue); the fp32 parity
// path keeps erff.
__device__ __forceinline__ float erf_fast(float z) {
    const float az = fabsf(z);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));
    float p = 1.061405429f;
    p = fmaf(p, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(az * az * -1.4426950408889634f));
    return copysignf(fmaf(-p, e, 1.0f), z);
}
__device__ __forceinline__ float gelu_erf_fast(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f)); }

// ---------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) and the epilogue GELU built on it.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma_f2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul_f2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add_f2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// erf-GELU for the 16-bit GEMM epilogue, two elements per call:  x * Phi(x) = x / (1 + 2^(x * q(x^2))), where x*q(x^2) is a
// degree-9 odd polynomial fitted to -log2(e) * logit(Phi(x)) (tools/fit_gelu.py).  Max |error| vs the exact erf form is
// 3.9e-6 absolute over all x, i.e. < 0.05 bf16 ulp wherever |gelu(x)| >= 0.01: the value rounded to bf16/fp16 equals the
// rounded exact value for all but a few per cent of the elements that sit next to a rounding boundary.  6 packed FP32 +
// 4 MUFU issue slots per pair instead of ~50 for two erff() calls, which made the c_fc GEMM epilogue-bound.
// The fp32 parity path (gemm_f32.cu) keeps erff.
__device__ __forceinline__ void gelu_pair_fast(float& x0, float& x1) {
    constexpr float kL = -1.4426950408889634f;
    const uint64_t x = pack_f2(x0, x1);
    const uint64_t s = mul_f2(x, x);
    uint64_t q = fma_f2(pack_f2(2.28182475e-06f * kL, 2.28182475e-06f * kL), s, pack_f2(-6.19073477e-05f * kL, -6.19073477e-05f * kL));
    q = fma_f2(q, s, pack_f2(-2.45941020e-04f * kL, -2.45941020e-04f * kL));
    q = fma_f2(q, s, pack_f2(7.29314157e-02f * kL, 7.29314157e-02f * kL));
    q = fma_f2(q, s, pack_f2(1.59565838e+00f * kL, 1.59565838e+00f * kL));
    float u0, u1;
    unpack_f2(mul_f2(q, x), u0, u1);
    float d0, d1;
    unpack_f2(add_f2(pack_f2(ex2_fast(u0), ex2_fast(u1)), pack_f2(1.0f, 1.0f)), d0, d1);
    unpack_f2(mul_f2(x, pack_f2(rcp_fast(d0), rcp_fast(d1))), x0, x1);
}

// d/dx of the erf-GELU, two elements per call:  Phi(x) + x phi(x)  with the same fitted Phi as gelu_pair_fast and
// phi(x) = 2^(-x^2 log2(e) / 2) / sqrt(2 pi).  |error| < 1e-5 (the gradients it scales are 16-bit).
__device__ __forceinline__ void gelu_grad_pair_fast(float x0, float x1, float& d0, float& d1) {
    constexpr float kL = -1.4426950408889634f;
    const uint64_t x = pack_f2(x0, x1);
    const uint64_t s = mul_f2(x, x);
    uint64_t q = fma_f2(pack_f2(2.28182475e-06f * kL, 2.28182475e-06f * kL), s, pack_f2(-6.19073477e-05f * kL, -6.19073477e-05f * kL));
    q = fma_f2(q, s, pack_f2(-2.45941020e-04f * kL, -2.45941020e-04f * kL));
    q = fma_f2(q, s, pack_f2(7.29314157e-02f * kL, 7.29314157e-02f * kL));
    q = fma_f2(q, s, pack_f2(1.59565838e+00f * kL, 1.59565838e+00f * kL));
    float u0, u1, h0, h1;
    unpack_f2(mul_f2(q, x), u0, u1);
    unpack_f2(mul_f2(s, pack_f2(0.5f * kL, 0.5f * kL)), h0, h1);
    float e0, e1;
    unpack_f2(add_f2(pack_f2(ex2_fast(u0), ex2_fast(u1)), pack_f2(1.0f, 1.0f)), e0, e1);
    const uint64_t cdf = pack_f2(rcp_fast(e0), rcp_fast(e1));
    const uint64_t xpdf = mul_f2(x, pack_f2(0.3989422804014327f * ex2_fast(h0), 0.3989422804014327f * ex2_fast(h1)));
    unpack_f2(add_f2(cdf, xpdf), d0, d1);
}

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start while
// its predecessor in the stream is still draining; everything before pdl_wait() (barrier init, TMEM allocation, descriptor
// prefetch, shared-memory zero fill) overlaps the predecessor's tail, pdl_wait() blocks until the predecessor has completed
// and its writes are visible.  pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as soon as this kernel's
// CTAs free their SMs.  Both are no-ops when the launch did not use the attribute.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// Warp helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (sticky CUDA error the host reports) instead of a
// hung GPU.  ~2^31 cycles is > 1 s at any B200 clock; a healthy wait is microseconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 31)) {
            printf("b200clip: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

// ---------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2D tile load global -> shared, completion on an mbarrier
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
constexpr uint64_t kCacheHintEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kCacheHintEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kCacheHintEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}

// 2D tile prefetch global -> L2 (no shared memory, no completion tracking): pulls DRAM latency out of the operand ring
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* desc, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(c0),
                 "r"(c1)
                 : "memory");
}

// 2D tile store shared -> global (bulk async-group completion); out-of-bounds parts of the box are clipped
__device__ __forceinline__ void tma_store_2d(const void* desc, const void* smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* desc, const void* smem_src, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// same, but global[box] += shared[box] element-wise (performed by the L2 in the tensor map's element type)
__device__ __forceinline__ void tma_reduce_add_2d(const void* desc, const void* smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM -> register loads
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 and fp16 inputs with fp32 accumulation.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All previously issued tcgen05.mma of this thread arrive (once) on `bar` when they complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (warp%4)*32+t.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM, same shape: thread t of the warp writes lane (warp%4)*32+t, 32 consecutive columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of a 2-cluster drive one 256-row UMMA together
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 remAddr32;\n\t"
        "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t"
        "}\n"
        :
        : "r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
// In a CTA pair the peer's copy of a shared-memory object differs in address bit 24; clearing it addresses the
// even (leader) CTA's copy from either CTA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (both CTAs of the pair issue it)
__device__ __forceinline__ void tma_load_2d_pair(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1,
                                                 uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
          "l"(hint)
        : "memory");
}
// 4-D tile load (NHWC pixel block, see make_tmap_nhwc), same completion mechanism
__device__ __forceinline__ void tma_load_4d_pair(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1, int32_t c2,
                                                 int32_t c3, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2),
          "r"(c3), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1, int32_t c2,
                                                 int32_t c3, int32_t c4, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2),
          "r"(c3), "r"(c4), "l"(hint)
        : "memory");
}
// Same, multicast: the box lands at the same CTA-relative offset in every CTA of `cta_mask`, and each destination's bytes are
// credited to the mbarrier at this offset in the leader (even) CTA of that destination's pair.
__device__ __forceinline__ void tma_load_2d_pair_mcast(const void* desc, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1,
                                                       uint16_t cta_mask, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
        " [%0], [%1, {%4, %5}], [%2], %3, %6;"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & kPeerBitMask), "h"(cta_mask), "r"(c0),
          "r"(c1), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit arrives on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

// Shared-memory matrix descriptor for a K-major operand tile written by TMA with SWIZZLE_128B
// (row pitch 128 B, 8-row / 1024 B swizzle atoms).  Fields (PTX ISA "matrix descriptor"):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4 (unused for SW128 K-major)
//   [32,46) stride-dim byte offset >> 4 (= 1024 B between 8-row groups)
//   [46,48) version = 1 (sm_100)      [61,64) layout: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

// Shared-memory descriptor of an MN-major operand tile written by TMA with SWIZZLE_128B as [K rows][64 MN elements] boxes (128 B
// rows, 8-row / 1024 B swizzle atoms), the 64-element MN blocks `lbo_bytes` apart.  Canonical form
// ((64,m),(8,k)):((1,LBO),(64,SBO)) in elements: SBO = 1024 B between 8-row K groups, LBO = distance between MN blocks.  One
// UMMA K step (16 rows) advances the start address by 2048 B.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024u >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

// K-major operand tile with a narrower swizzle (rows of 64 B: layout 4 = SWIZZLE_64B, 8-row groups 512 B apart; rows of 32 B:
// layout 6 = SWIZZLE_32B, groups 256 B apart) -- the patch-embedding A tiles, whose rows are one pixel row of a patch
__device__ __forceinline__ uint64_t make_narrow_kmajor_desc(uint32_t smem_addr, uint32_t row_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((8u * row_bytes) >> 4) << 32;
    d |= 1ull << 46;
    d |= static_cast<uint64_t>(row_bytes == 64 ? 4u : 6u) << 61;
    return d;
}

// Instruction descriptor for kind::f16: fp32 accumulate, A/B both K-major, no negate/saturate/sparsity.
//   [4,6) c_format = 1 (F32)  [7,10) a_format  [10,13) b_format  [15] a_major  [16] b_major
//   [17,23) N >> 3            [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t fmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------
// Vector global access
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg128(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

}  // namespace b200clip
