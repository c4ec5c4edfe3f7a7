// Host-side TMA tensor-map construction shared by the tcgen05 GEMM kernels.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace b200clip {

// 2D row-major [rows, cols] 16-bit matrix with row pitch `ld` elements; box = [box_rows, box_cols] with box_cols * 2 B
// == 128 B (one SWIZZLE_128B row).  Out-of-bounds box elements read as zero and are clipped on stores.
int make_tmap_2d(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols);

}  // namespace b200clip
