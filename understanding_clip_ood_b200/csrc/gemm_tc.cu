// tcgen05 / TMEM / TMA GEMM for the CLIP towers:  C[M,N] = epilogue(A[M,K] · W[N,K]^T + bias).
//
// B200-native design (not a port: the reference only calls F.linear, transformer.py:224-263):
//   * persistent kernel, one CTA per SM, 12 warps with fixed roles:
//       warp 0   TMA producer (one thread): A / W tiles -> 128B-swizzled shared memory ring
//       warp 1   MMA issuer (one thread): tcgen05.mma kind::f16, 128 x BLOCK_N x 16 per instruction
//       warp 2   TMEM allocator / deallocator
//       warp 4..11  epilogue: tcgen05.ld accumulator -> registers -> bias / GELU / residual -> global
//   * accumulators live in TMEM, double buffered (2 x BLOCK_N columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1;
//   * three mbarrier pipelines: smem full/empty (TMA <-> MMA), tmem full/empty (MMA <-> epilogue);
//   * tiles are walked in groups of `group_m` row-tiles so concurrently running CTAs share W and A
//     tiles in the 126 MB L2.
// Rounding points mirror the reference's bf16/fp16 eager path: the linear output (acc + bias) is
// rounded to the storage type before the activation / residual add, which round again.
#include "common.cuh"
#include "tmap.h"

#include <mutex>

namespace b200clip {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 x 16-bit = one 128 B swizzle row
constexpr int kUmmaK = 16;
constexpr int kAccStages = 2;
constexpr int kNumThreads = 384;
constexpr int kEpilogueWarp0 = 4;
constexpr int kNumEpilogueWarps = 8;

struct GemmParams {
    const void* bias;
    const void* residual;
    const float* pos;
    void* C;
    int M, N, K;
    int64_t ldc, ldr;
    int g_in, g_out;
    int group_m;
    int m_tiles, n_tiles;
};

template <int BLOCK_N> struct TileCfg {
    static constexpr int kStages = BLOCK_N == 256 ? 4 : 6;
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarrierBytes = (2 * kStages + 2 * kAccStages) * 8 + 16;
    static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;  // +1024: manual alignment
    static constexpr int kTmemCols = kAccStages * BLOCK_N;
};

__device__ __forceinline__ void tile_coords(int t, const GemmParams& p, int& mt, int& nt) {
    const int per_group = p.group_m * p.n_tiles;
    const int g = t / per_group;
    const int r = t - g * per_group;
    const int m0 = g * p.group_m;
    const int gm = min(p.group_m, p.m_tiles - m0);
    nt = r / gm;
    mt = m0 + (r - nt * gm);
}

template <typename T, int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
    using Cfg = TileCfg<BLOCK_N>;
    using H = Half16<T>;
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024 B alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + kAccStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.n_tiles;
    const int num_kb = (p.K + kBlockK - 1) / kBlockK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], kNumEpilogueWarps);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int mt, nt;
                tile_coords(t, p, mt, nt);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    // activations stream through once per N-tile: default policy; weights are re-read by every
                    // M-tile: keep them in L2
                    tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * Cfg::kABytes, kb * kBlockK, mt * kBlockM,
                                kCacheHintEvictNormal);
                    tma_load_2d(&tmap_w, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBlockK, nt * BLOCK_N,
                                kCacheHintEvictLast);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(H::kUmmaFormat, kBlockM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t desc_a = make_sw128_kmajor_desc(smem_u32(smem_a + stage * Cfg::kABytes));
                    const uint64_t desc_b = make_sw128_kmajor_desc(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // advance 16 elements = 32 B along K inside the 128 B swizzle row: +2 in the (>>4) address field
                        umma_f16(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                    if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= kEpilogueWarp0) {
        // ===================== epilogue =====================
        const int e = warp - kEpilogueWarp0;
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int col_half = e >> 2;            // which half of the BLOCK_N columns
        constexpr int kColsPerWarp = BLOCK_N / 2;
        const T* bias = static_cast<const T*>(p.bias);
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            int mt, nt;
            tile_coords(t, p, mt, nt);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();

            const int row = mt * kBlockM + q * 32 + lane;
            const bool row_ok = row < p.M;
            int64_t out_row = row;
            const float* pos_row = nullptr;
            if constexpr (EPI == 4) {
                const int img = row / p.g_in;
                const int pi = row - img * p.g_in;
                out_row = static_cast<int64_t>(img) * p.g_out + pi + 1;
                pos_row = p.pos + static_cast<int64_t>(pi + 1) * p.N;
            }
            T* c_row = static_cast<T*>(p.C) + out_row * p.ldc;
            const T* r_row = nullptr;
            if constexpr (EPI == 3) r_row = static_cast<const T*>(p.residual) + static_cast<int64_t>(row) * p.ldr;

#pragma unroll 1
            for (int c = 0; c < kColsPerWarp; c += 32) {
                const int col_in_tile = col_half * kColsPerWarp + c;
                const int col0 = nt * BLOCK_N + col_in_tile;
                uint32_t v[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + col_in_tile;
                tmem_ld_32x32(taddr, v);

                // operands of the epilogue are fetched while the TMEM load is in flight
                uint4 bvec[4];
                uint4 rvec[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int col = col0 + g * 8;
                    bvec[g] = make_uint4(0, 0, 0, 0);
                    rvec[g] = make_uint4(0, 0, 0, 0);
                    if (col < p.N) {
                        if (bias != nullptr) bvec[g] = ldg128(bias + col);
                        if constexpr (EPI == 3) {
                            if (row_ok) rvec[g] = *reinterpret_cast<const uint4*>(r_row + col);
                        }
                    }
                }
                tmem_ld_wait();
                if (c + 32 >= kColsPerWarp) {
                    // last TMEM read of this tile: hand the accumulator stage back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                }

#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int col = col0 + g * 8;
                    const uint32_t bw[4] = {bvec[g].x, bvec[g].y, bvec[g].z, bvec[g].w};
                    const uint32_t rw[4] = {rvec[g].x, rvec[g].y, rvec[g].z, rvec[g].w};
                    uint32_t ow[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 b2 = H::unpack(bw[j]);
                        float x0 = round16<T>(__uint_as_float(v[g * 8 + 2 * j]) + b2.x);
                        float x1 = round16<T>(__uint_as_float(v[g * 8 + 2 * j + 1]) + b2.y);
                        if constexpr (EPI == 1) {
                            x0 = gelu_erf(x0);
                            x1 = gelu_erf(x1);
                        } else if constexpr (EPI == 2) {
                            x0 = quick_gelu(x0);
                            x1 = quick_gelu(x1);
                        } else if constexpr (EPI == 3) {
                            const float2 r2 = H::unpack(rw[j]);
                            x0 += r2.x;
                            x1 += r2.y;
                        } else if constexpr (EPI == 4) {
                            if (row_ok && col < p.N) {
                                x0 += round16<T>(pos_row[col + 2 * j]);
                                x1 += round16<T>(pos_row[col + 2 * j + 1]);
                            }
                        }
                        ow[j] = H::pack(x0, x1);
                    }
                    if (row_ok && col < p.N) stg128(c_row + col, make_uint4(ow[0], ow[1], ow[2], ow[3]));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
// 2D row-major [rows, cols] 16-bit matrix, box = [box_rows, 64 cols], 128B swizzle, OOB reads give zeros.
int make_tmap(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    return make_tmap_2d(map, is_bf16, ptr, rows, cols, ld, box_rows, kBlockK);
}

template <typename T, int BLOCK_N, int EPI>
int launch_one(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, cudaStream_t stream) {
    using Cfg = TileCfg<BLOCK_N>;
    auto kern = gemm_tc_kernel<T, BLOCK_N, EPI>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(gemm_tc smem)");
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(ta, tw, p);
    B2C_LAUNCH_CHECK("gemm_tc_kernel");
    return 0;
}

template <typename T, int BLOCK_N>
int launch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, cudaStream_t s) {
    switch (epi) {
        case 0: return launch_one<T, BLOCK_N, 0>(ta, tw, p, s);
        case 1: return launch_one<T, BLOCK_N, 1>(ta, tw, p, s);
        case 2: return launch_one<T, BLOCK_N, 2>(ta, tw, p, s);
        case 3: return launch_one<T, BLOCK_N, 3>(ta, tw, p, s);
        case 4: return launch_one<T, BLOCK_N, 4>(ta, tw, p, s);
    }
    set_last_error("gemm: unknown epilogue %d", epi);
    return -1;
}

// Pick the N tile that wastes the least of the last wave (M=6400-class problems are only a few waves).
int pick_block_n(int M, int N) {
    const int sms = num_sms();
    const int mt = (M + kBlockM - 1) / kBlockM;
    double best_cost = 1e30;
    int best = 256;
    const int cands[2] = {256, 128};
    for (int bn : cands) {
        const int nt = (N + bn - 1) / bn;
        const long tiles = static_cast<long>(mt) * nt;
        const long waves = (tiles + sms - 1) / sms;
        // cost in "tile-columns": waves * bn, with a small penalty for the narrower tile (more A re-reads / epilogue overhead)
        const double cost = static_cast<double>(waves) * bn * (bn == 128 ? 1.06 : 1.0);
        if (cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

}  // namespace

int gemm_tc(bool is_bf16, const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, const void* residual,
            int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int epilogue, const float* pos, int g_in, int g_out,
            int force_block_n, cudaStream_t stream) {
    B2C_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
    B2C_CHECK_ARG(N % 8 == 0, "gemm: N=%d must be a multiple of 8", N);
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
    if (epilogue == 4) {
        B2C_CHECK_ARG(pos != nullptr && g_in > 0 && g_out == g_in + 1, "gemm: patch epilogue needs pos, g_in, g_out=g_in+1");
    }
    const int bn = force_block_n > 0 ? force_block_n : pick_block_n(M, N);
    B2C_CHECK_ARG(bn == 128 || bn == 256, "gemm: BLOCK_N must be 128 or 256");

    CUtensorMap ta, tw;
    if (make_tmap(&ta, is_bf16, A, M, K, lda, kBlockM) != 0) return -1;
    if (make_tmap(&tw, is_bf16, W, N, K, ldw, bn) != 0) return -1;

    GemmParams p;
    p.bias = bias;
    p.residual = residual;
    p.pos = pos;
    p.C = C;
    p.M = M;
    p.N = N;
    p.K = K;
    p.ldc = ldc;
    p.ldr = ldr;
    p.g_in = g_in;
    p.g_out = g_out;
    p.group_m = 16;
    p.m_tiles = (M + kBlockM - 1) / kBlockM;
    p.n_tiles = (N + bn - 1) / bn;

    if (is_bf16) {
        return bn == 256 ? launch_epi<__nv_bfloat16, 256>(epilogue, ta, tw, p, stream)
                         : launch_epi<__nv_bfloat16, 128>(epilogue, ta, tw, p, stream);
    }
    return bn == 256 ? launch_epi<__half, 256>(epilogue, ta, tw, p, stream)
                     : launch_epi<__half, 128>(epilogue, ta, tw, p, stream);
}

}  // namespace b200clip
