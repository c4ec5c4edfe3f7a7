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

// The same 4-D box over a tensor with explicit element pitches of the W, H and B dimensions (C contiguous): e.g. the token rows
// 1 .. G*G of x [B, G*G + 1, width] seen as [B, G, G, width].
int make_tmap_nhwc_strided(CUtensorMap* map, bool is_bf16, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint64_t pitch_w,
                           uint64_t pitch_h, uint64_t pitch_b, uint32_t bb, uint32_t by, uint32_t bx);

// 5D view of an NCHW image batch [B, 3, S, S] (16-bit) for the implicit patch-embedding GEMM: dimensions (dx within a patch, dy
// within a patch, patch column, patch row, channel + 3 * image), box = (P, 1, bx, by, bb images at ONE channel through an element
// stride of 3 on the last dimension): bx * by * bb = 128 patches x P elements = a K-major operand sub-tile whose rows are ONE pixel
// row of each patch, P * 2 bytes = one swizzle row (SWIZZLE_64B for P = 32, SWIZZLE_32B for P = 16); 64 / P such boxes make a K block.
int make_tmap_patches(CUtensorMap* map, bool is_bf16, const void* image, uint64_t B, uint64_t S, uint32_t P, uint32_t rows, uint32_t bb,
                      uint32_t by, uint32_t bx);

}  // namespace b200clip
