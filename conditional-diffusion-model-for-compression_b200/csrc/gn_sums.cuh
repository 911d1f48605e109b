// GroupNorm statistics as fixed-point integer accumulators.
//
// Every producer (conv epilogue, gn_stats_kernel) adds its per-tile / per-strip (sum, sum of squares) of each
// (image, group) with ONE 64-bit integer atomic per value: integer addition is associative, so the totals are
// bitwise reproducible whatever the arrival order, and no partial buffer / finalize pass is needed -- the consumer
// (gn_apply_kernel) reads 2 x 8 bytes per group and derives y = a*x + b itself.  The decode plan gives every
// GroupNorm of a step its own slot and clears all slots with one memset node at the start of the step.
//
// Scale 2^20: a partial (fp32, 24-bit mantissa) loses at most 2^-21 absolute; totals stay far below 2^63
// (|sum of squares| of a level-0 group at 2048^2 < 2^33 before scaling).
// Oracle counterpart: oracle/unet.py RB / Attn GroupNorm (the reference ships no code).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cdc {

typedef long long gn_sum_t;                    // slot layout: [image][32 groups][2]
constexpr float kGnFixScale = 1048576.0f;      // 2^20
constexpr double kGnFixInv = 1.0 / 1048576.0;

__device__ __forceinline__ void gn_sums_add(gn_sum_t* slot_of_image, int group, float s, float q) {
    unsigned long long* d = reinterpret_cast<unsigned long long*>(slot_of_image + group * 2);
    atomicAdd(d, static_cast<unsigned long long>(__float2ll_rn(s * kGnFixScale)));      // two's complement: signed sums add up
    atomicAdd(d + 1, static_cast<unsigned long long>(__float2ll_rn(q * kGnFixScale)));
}

}  // namespace cdc
