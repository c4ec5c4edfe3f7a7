// fp32 GEMM (the 1e-4 / index-exact parity mode and the ClipLoss contractions):
//   C[M,N] = epilogue(op(A) · op(B) + bias) with true fp32 FFMA accumulation.
// TF32 tensor cores would break the fp32 parity gate (SURVEY.md §7 hard part 1: logit error must stay ~1e-7), so this
// path deliberately stays on the FP32 pipe.  128x128x16 tiles, 256 threads, 8x8 outputs per thread, operands staged
// k-major in shared memory.  op(A): A[M,K] row-major (TA=0) or A stored as [K,M] (TA=1);
// op(B): nn.Linear layout W[N,K] (TB=0) or B stored as [K,N] (TB=1).
#include "common.cuh"
#include "internal.h"

namespace b200clip {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;

struct F32Params {
    const float* A;
    const float* W;
    const float* bias;
    const float* residual;
    const float* pos;
    float* C;
    int M, N, K;
    int64_t lda, ldw, ldc, ldr;
    int g_in, g_out;
};

// stage a [BK x 128] k-major tile of an operand stored with the contraction index contiguous ([rows, K])
__device__ __forceinline__ void stage_kcontig(float (*S)[BM + 4], const float* base, int64_t ld, int r0, int rows, int k0, int K,
                                              int tid) {
    const int lrow = tid >> 2;
    const int lk = (tid & 3) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = lrow + h * 64;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < rows && k0 + lk < K) a = *reinterpret_cast<const float4*>(base + (int64_t)(r0 + r) * ld + k0 + lk);
        S[lk + 0][r] = a.x; S[lk + 1][r] = a.y; S[lk + 2][r] = a.z; S[lk + 3][r] = a.w;
    }
}
// ... or stored with the output index contiguous ([K, rows])
__device__ __forceinline__ void stage_mcontig(float (*S)[BM + 4], const float* base, int64_t ld, int r0, int rows, int k0, int K,
                                              int tid) {
    const int r = (tid & 31) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = (tid >> 5) + h * 8;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + k < K && r0 + r < rows) a = *reinterpret_cast<const float4*>(base + (int64_t)(k0 + k) * ld + r0 + r);
        *reinterpret_cast<float4*>(&S[k][r]) = a;
    }
}

template <int EPI, bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const F32Params p) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int tx = tid & 15;   // column group
    const int ty = tid >> 4;   // row group

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += BK) {
        if constexpr (TA) stage_mcontig(As, p.A, p.lda, m0, p.M, k0, p.K, tid);
        else stage_kcontig(As, p.A, p.lda, m0, p.M, k0, p.K, tid);
        if constexpr (TB) stage_mcontig(Ws, p.W, p.ldw, n0, p.N, k0, p.K, tid);
        else stage_kcontig(Ws, p.W, p.ldw, n0, p.N, k0, p.K, tid);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], w[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= p.M) continue;
        int64_t out_row = row;
        const float* pos_row = nullptr;
        if constexpr (EPI == 4) {
            const int img = row / p.g_in;
            const int pi = row - img * p.g_in;
            out_row = (int64_t)img * p.g_out + pi + 1;
            pos_row = p.pos + (int64_t)(pi + 1) * p.N;
        }
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int col = n0 + jh * 64 + tx * 4;
            if (col >= p.N) continue;  // N % 4 == 0 is checked on the host
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][jh * 4 + j];
                if (p.bias != nullptr) x += p.bias[col + j];
                if constexpr (EPI == 1) x = gelu_erf(x);
                else if constexpr (EPI == 2) x = x / (1.0f + expf(-1.702f * x));
                else if constexpr (EPI == 3 || EPI == 6) x += p.residual[(int64_t)row * p.ldr + col + j];
                else if constexpr (EPI == 4) x += pos_row[col + j];
                if constexpr (EPI == 5 || EPI == 6) x = fmaxf(x, 0.f);
                o[j] = x;
            }
            *reinterpret_cast<float4*>(p.C + out_row * p.ldc + col) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

}  // namespace

int gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* residual, int64_t ldr,
             float* C, int64_t ldc, int M, int N, int K, int epilogue, const float* pos, int g_in, int g_out,
             cudaStream_t stream) {
    B2C_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_f32: empty problem M=%d N=%d K=%d", M, N, K);
    B2C_CHECK_ARG(N % 4 == 0 && K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0,
                  "gemm_f32: N, K and leading dimensions must be multiples of 4");
    B2C_CHECK_ARG((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) % 16 == 0,
                  "gemm_f32: A, W, C must be 16-byte aligned");
    if (epilogue == 3 || epilogue == 6) B2C_CHECK_ARG(residual != nullptr, "gemm_f32: residual epilogue needs a residual pointer");
    if (epilogue == 4) B2C_CHECK_ARG(pos != nullptr && g_in > 0 && g_out == g_in + 1, "gemm_f32: bad patch epilogue arguments");
    F32Params p{A, W, bias, residual, pos, C, M, N, K, lda, ldw, ldc, ldr, g_in, g_out};
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    switch (epilogue) {
        case 0: gemm_f32_kernel<0, false, false><<<grid, 256, 0, stream>>>(p); break;
        case 1: gemm_f32_kernel<1, false, false><<<grid, 256, 0, stream>>>(p); break;
        case 2: gemm_f32_kernel<2, false, false><<<grid, 256, 0, stream>>>(p); break;
        case 3: gemm_f32_kernel<3, false, false><<<grid, 256, 0, stream>>>(p); break;
        case 4: gemm_f32_kernel<4, false, false><<<grid, 256, 0, stream>>>(p); break;
        case 5: gemm_f32_kernel<5, false, false><<<grid, 256, 0, stream>>>(p); break;
        case 6: gemm_f32_kernel<6, false, false><<<grid, 256, 0, stream>>>(p); break;
        default: set_last_error("gemm_f32: unknown epilogue %d", epilogue); return -1;
    }
    B2C_LAUNCH_CHECK("gemm_f32_kernel");
    return 0;
}

// C[M,N] = op(A) op(B) without epilogue; ta: A stored [K,M]; tb: B stored [K,N] (tb = 0: B stored [N,K]).
int gemm_f32_nt(bool ta, bool tb, const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc, int M, int N,
                int K, cudaStream_t stream) {
    B2C_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_f32_nt: empty problem");
    B2C_CHECK_ARG(N % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 && (ta ? M % 4 == 0 : K % 4 == 0) &&
                      (tb ? true : K % 4 == 0),
                  "gemm_f32_nt: dimensions must be multiples of 4");
    F32Params p{A, Bm, nullptr, nullptr, nullptr, C, M, N, K, lda, ldb, ldc, 0, 0, 0};
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    if (!ta && !tb) gemm_f32_kernel<0, false, false><<<grid, 256, 0, stream>>>(p);
    else if (!ta && tb) gemm_f32_kernel<0, false, true><<<grid, 256, 0, stream>>>(p);
    else if (ta && !tb) gemm_f32_kernel<0, true, false><<<grid, 256, 0, stream>>>(p);
    else gemm_f32_kernel<0, true, true><<<grid, 256, 0, stream>>>(p);
    B2C_LAUNCH_CHECK("gemm_f32_kernel(nt)");
    return 0;
}

}  // namespace b200clip
