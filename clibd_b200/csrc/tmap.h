// Host-side creation of TMA tensor maps (cuTensorMapEncodeTiled, fetched through the runtime's
// driver-entry-point API so the library does not link against libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace clibd {

// 2-D row-major tensor of 16-bit elements [rows, cols] with row pitch `pitch_elems`;
// box = {box_cols (must be 64 => 128 B, SWIZZLE_128B), box_rows <= 256}; out-of-bounds -> zeros.
// Returns 0 on success (error text via set_error).
int make_tmap_2d_16bit(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                       uint32_t box_cols, uint32_t box_rows, int fmt_bf16);

}  // namespace clibd
