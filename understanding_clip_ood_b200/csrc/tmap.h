// Host-side TMA tensor-map construction shared by the tcgen05 GEMM kernels.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace b200clip {

// 2D row-major [rows, cols] 16-bit matrix with row pitch `ld` elements; box = [box_rows, box_cols] with box_cols * 2 B
// == 128 B (one SWIZZLE_128B row).  Out-of-bounds box elements read as zero and are clipped on stores.
int make_tmap_2d(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols);

// 4D NHWC activation tensor [B, H, W, C] (16-bit, densely packed) with a box of [bb images, by rows, bx pixels, 64 channels]: the
// box lands in shared memory as bb * by * bx rows of 128 B (SWIZZLE_128B), i.e. a K-major UMMA operand tile whose rows are the
// pixels of a small image block.  Coordinates may lie outside the tensor (negative, or beyond W / H / B): such elements read as
// zero -- the zero padding of a convolution -- and are clipped on stores.  Not memoised (built once per convolution call).
int make_tmap_nhwc(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t bb, uint32_t by,
                   uint32_t bx);

}  // namespace b200clip
