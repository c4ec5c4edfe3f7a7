// cuTensorMapEncodeTiled is fetched from the driver at run time (no link-time dependency on libcuda).
#include "tmap.h"

#include <cuda_runtime.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace b200clip {

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// Encoding a tensor map costs several microseconds of host time and the towers issue ~200 per forward with operands
// (weights, workspace slices) that repeat from call to call: memoise by the full argument tuple.  The descriptor is a
// pure function of the key, so a hit can never be stale.
struct TmapKey {
    const void* ptr;
    uint64_t rows, cols, ld;
    uint32_t box_rows, box_cols, is_bf16;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols &&
               is_bf16 == o.is_bf16;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = reinterpret_cast<uintptr_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
        auto mix = [&h](uint64_t v) { h = (h ^ v) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; };
        mix(k.rows); mix(k.cols); mix(k.ld); mix((uint64_t(k.box_rows) << 32) | (uint64_t(k.box_cols) << 1) | k.is_bf16);
        return static_cast<size_t>(h);
    }
};
constexpr size_t kTmapCacheMax = 8192;
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>& tmap_cache() {
    static auto* m = new std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>();
    return *m;
}

}  // namespace

int make_tmap_2d(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols) {
    const TmapKey key{ptr, rows, cols, ld, box_rows, box_cols, is_bf16 ? 1u : 0u};
    {
        std::lock_guard<std::mutex> lk(g_tmap_mu);
        auto& c = tmap_cache();
        auto it = c.find(key);
        if (it != c.end()) {
            *map = it->second;
            return 0;
        }
    }
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        set_last_error("cuTensorMapEncodeTiled not available from the driver");
        return -1;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * 2};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (CUresult %d) for ptr=%p rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r, ptr,
                       (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
        return -1;
    }
    {
        std::lock_guard<std::mutex> lk(g_tmap_mu);
        auto& c = tmap_cache();
        if (c.size() >= kTmapCacheMax) c.clear();
        c.emplace(key, *map);
    }
    return 0;
}

int make_tmap_nhwc_strided(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint64_t pitch_w,
                           uint64_t pitch_h, uint64_t pitch_b, uint32_t bb, uint32_t by, uint32_t bx) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        set_last_error("cuTensorMapEncodeTiled not available from the driver");
        return -1;
    }
    const cuuint64_t dims[4] = {C, W, H, B};
    const cuuint64_t strides[3] = {pitch_w * 2, pitch_h * 2, pitch_b * 2};
    const cuuint32_t box[4] = {64, bx, by, bb};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled (NHWC) failed (CUresult %d) for ptr=%p B=%llu H=%llu W=%llu C=%llu box=%ux%ux%u", (int)r, ptr,
                       (unsigned long long)B, (unsigned long long)H, (unsigned long long)W, (unsigned long long)C, bb, by, bx);
        return -1;
    }
    return 0;
}

int make_tmap_nhwc(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t bb, uint32_t by,
                   uint32_t bx) {
    return make_tmap_nhwc_strided(map, is_bf16, ptr, B, H, W, C, C, W * C, H * W * C, bb, by, bx);
}

int make_tmap_patches(CUtensorMap* map, bool is_bf16, const void* image, uint64_t B, uint64_t S, uint32_t P, uint32_t rows, uint32_t bb,
                      uint32_t by, uint32_t bx) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        set_last_error("cuTensorMapEncodeTiled not available from the driver");
        return -1;
    }
    const uint64_t G = S / P;
    (void)rows;   // one pixel row of every patch per box: the inner box extent (P pixels) is exactly one swizzle row
    const cuuint64_t dims[5] = {P, P, G, G, 3 * B};
    const cuuint64_t strides[4] = {S * 2, static_cast<cuuint64_t>(P) * 2, S * P * 2, S * S * 2};
    const cuuint32_t box[5] = {P, 1, bx, by, bb * 3};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 3};
    const CUtensorMapSwizzle swz = P * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(image), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled (patches) failed (CUresult %d) for image=%p B=%llu S=%llu P=%u box=%ux%ux%ux%u", (int)r, image,
                       (unsigned long long)B, (unsigned long long)S, P, rows, bx, by, bb);
        return -1;
    }
    return 0;
}

}  // namespace b200clip
