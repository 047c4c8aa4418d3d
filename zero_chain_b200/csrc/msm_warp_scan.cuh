// Warp-level products for the batched-affine rounds (msm_batchaff.cuh): every thread of a warp holds the product T_lane of its
// denominators; the backward pass needs 1 / T_lane, and the round has ONE inversion.  Two Kogge-Stone scans over the warp give each
// lane the product of the OTHER 31 totals and the warp's total, so 1 / T_lane = (1 / total) * others.
// Kept apart from the kernels (no cp.async, no launch bounds) so that tests/host_emul can compile it for the host with a SIMT shim.
#pragma once
#include <stdint.h>
#include "curve.cuh"

namespace zkmsm {

// ---- warp-level products of the thread totals (no block barriers: the warps of a block stay independent) ----
template <class F>
__device__ __forceinline__ F ba_shfl(const F &v, int delta, bool up) {
    F r;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&v);
    uint32_t *d = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 4); k++) d[k] = up ? __shfl_up_sync(0xffffffffu, s[k], delta) : __shfl_down_sync(0xffffffffu, s[k], delta);
    return r;
}
template <class F>
__device__ __forceinline__ F ba_sel(bool c, const F &a, const F &b) {
    F r;
    const uint32_t *pa = reinterpret_cast<const uint32_t *>(&a), *pb = reinterpret_cast<const uint32_t *>(&b);
    uint32_t *d = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 4); k++) d[k] = c ? pa[k] : pb[k];
    return r;
}
// For the 32 thread totals T_0 .. T_31 of a warp: `others` = product of all T_j, j != lane (prefix x suffix, two Kogge-Stone scans
// of 5 steps each), `all` = the warp's total in every lane.  1 / T_lane = (1 / all) * others.
template <class F>
__device__ __forceinline__ void ba_warp_products(const F &t, F &others, F &all) {
    const int lane = threadIdx.x & 31;
    F inc = t, dec = t;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        F y = ba_shfl(inc, d, true), p = inc * y;
        inc = ba_sel(lane >= d, p, inc);
        F z = ba_shfl(dec, d, false), q = dec * z;
        dec = ba_sel(lane + d < 32, q, dec);
    }
    F pre = ba_shfl(inc, 1, true), suf = ba_shfl(dec, 1, false);
    pre = ba_sel(lane == 0, F::one(), pre);
    suf = ba_sel(lane == 31, F::one(), suf);
    others = pre * suf;
    F tot;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&inc);
    uint32_t *dd = reinterpret_cast<uint32_t *>(&tot);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 4); k++) dd[k] = __shfl_sync(0xffffffffu, s[k], 31);
    all = tot;
}

}  // namespace zkmsm
