// Host-side runtime helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace ld {
void count_launch(int n = 1);
// rank-4 bf16 tensor map, SWIZZLE_128B, zero OOB fill. dims = {inner, outer, batch2, batch1} (elements),
// strides_bytes = byte strides of dims 1..3 (multiples of 16).
int encode_tmap_bf16_4d(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                        uint32_t box_inner, uint32_t box_outer, int swizzle_bytes = 128);
int encode_tmap_bf16_4d_box(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                            const uint32_t box[4], const uint32_t estr[4], int swizzle_bytes = 128);
}  // namespace ld
