// GroupNorm statistics as fixed-point integer accumulators.
//
// Every producer (conv epilogue, gn_stats_kernel) adds its per-tile / per-strip (sum, sum of squares) of each
// (image, group) with 64-bit integer atomics: integer addition is associative, so the totals are
// bitwise reproducible whatever the arrival order, and no partial buffer / finalize pass is needed -- the consumer
// (gn_apply_kernel, the APPLY warps of conv_kf.cu) reads 3 x 8 bytes per group and derives y = a*x + b itself.
// The decode plan gives every GroupNorm of a step its own slot and clears all slots with one memset node at the
// start of the step.
//
// Layout: [image][32 groups][4] = (sum, squares_lo, squares_hi, unused).
//   sum        : scale 2^20.  |partial| <= 2^29 (a 21-row strip of fp16-max values) -> 2^49 scaled; a 2048^2 image
//                has ~2^11 partials per group: total < 2^60.
//   squares    : split so that neither precision nor range depends on the activations' magnitude.  A partial
//                q (fp32, >= 0) is cut at 2^10: lo = q mod 1024 at scale 2^20 (< 2^30 per partial), hi = floor(q / 1024)
//                as a plain integer (< 2^35 even if every element is the fp16 maximum).  Both cuts are exact in fp32.
//                Round 1 kept q * 2^20 in ONE accumulator, which overflows int64 once a level-0 group's rms passes
//                ~1000 at 2048^2; trained weights may do that, synthetic ones do not.
// Every partial loses at most 2^-21 absolute, whatever its size.
// Oracle counterpart: oracle/unet.py RB / Attn GroupNorm (the reference ships no code).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cdc {

typedef long long gn_sum_t;
constexpr int kGnVals = 4;                     // values per (image, group)
constexpr int kGnImgStride = 32 * kGnVals;     // values per image
constexpr float kGnFixScale = 1048576.0f;      // 2^20
constexpr double kGnFixInv = 1.0 / 1048576.0;
constexpr float kGnSqCut = 1024.0f;            // 2^10

__device__ __forceinline__ void gn_sums_add(gn_sum_t* slot_of_image, int group, float s, float q) {
    unsigned long long* d = reinterpret_cast<unsigned long long*>(slot_of_image + group * kGnVals);
    atomicAdd(d, static_cast<unsigned long long>(__float2ll_rn(s * kGnFixScale)));      // two's complement: signed sums add up
    const float hi = floorf(q * (1.0f / kGnSqCut));
    const float lo = fmaf(-hi, kGnSqCut, q);                                             // exact: q has 24 significant bits
    atomicAdd(d + 1, static_cast<unsigned long long>(__float2ll_rn(lo * kGnFixScale)));
    if (hi != 0.0f) atomicAdd(d + 2, static_cast<unsigned long long>(__float2ll_rn(hi)));
}

// (sum, sum of squares) of one (image, group) as doubles (exact integer totals -> deterministic)
__device__ __forceinline__ void gn_sums_read(const gn_sum_t* a, double& s, double& q) {
    s = static_cast<double>(a[0]) * kGnFixInv;
    q = static_cast<double>(a[1]) * kGnFixInv + static_cast<double>(a[2]) * static_cast<double>(kGnSqCut);
}

}  // namespace cdc
