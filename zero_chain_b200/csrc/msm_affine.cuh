// Batched-affine pre-reduction of the MSM buckets (two levels of pairwise additions before the XYZZ
// accumulation; see msm.cuh for where it sits in the schedule).
//
// A mixed XYZZ addition costs 10 Montgomery products.  An AFFINE addition costs 1 inversion + 2M + 1S, and
// with Montgomery's simultaneous-inversion trick the inversion is shared: each thread takes AFF_B independent
// pair additions (consecutive outputs of the level), multiplies their denominators into a running product
// (1M per addition, prefix products parked in a global scratch array), inverts ONCE (binary extended Euclid,
// ~90 product-equivalents, all lanes of the warp running their own), and walks back (2M per addition to peel
// the individual inverses, then lambda = num * inv, x3 = lambda^2 - x0 - x1, y3 = lambda (x0 - x3) - y0).
// That is 6 products + 90/AFF_B per addition instead of 10.  Each level halves every bucket: bucket b with m
// points yields ceil(m/2) points (an odd leftover is copied), so the level's outputs are again grouped by
// bucket and offsets come from one scan.  Same group elements as bellman's bucket sums, so
// the canonical result cannot change; the exceptional cases the reference's add handles (P+P -> double,
// P+(-P) -> infinity, infinity operands; ec.rs:357-365, 394-397) are classified per pair below.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "curve.cuh"
#include "msm_affine_core.cuh"

namespace zkmsm {

constexpr int AFF_B = 64;          // pair additions per thread (one inversion each)

// sizes_out[b] = ceil(sizes_in[b] / 2)
static __global__ void k_half_sizes(const uint32_t *__restrict__ off_in, uint32_t *__restrict__ sizes_out, uint32_t n_buckets) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_buckets) sizes_out[b] = (off_in[b + 1] - off_in[b] + 1) >> 1;
}

// One level: out[o] for o in [t*AFF_B, (t+1)*AFF_B) — output o of bucket b is in[2j] + in[2j+1], j = o - off_out[b].
template <class F>
__global__ void __launch_bounds__(128) k_affine_round(const Affine<F> *__restrict__ in_pts, const uint32_t *__restrict__ sorted,
                                                      const uint32_t *__restrict__ off_in, const uint32_t *__restrict__ off_out,
                                                      uint32_t n_buckets, F *__restrict__ scratch, Affine<F> *__restrict__ out_pts) {
    const uint32_t total = off_out[n_buckets];
    const uint32_t o0 = (blockIdx.x * blockDim.x + threadIdx.x) * AFF_B;
    if (o0 >= total) return;
    const uint32_t o1 = o0 + AFF_B < total ? o0 + AFF_B : total;
    PairIO<F> io{in_pts, sorted};
    // bucket of o0: last b with off_out[b] <= o0
    uint32_t lo = 0, hi = n_buckets;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (off_out[mid] <= o0) lo = mid; else hi = mid; }
    const uint32_t b_first = lo;
    // ---- forward: running product of the denominators, prefix products to scratch ----
    F run = F::one();
    uint32_t b = b_first, b_end = off_out[b + 1], b_out0 = off_out[b], b_in0 = off_in[b], b_sz = off_in[b + 1] - b_in0;
    for (uint32_t o = o0; o < o1; o++) {
        while (o >= b_end) { b++; b_out0 = b_end; b_end = off_out[b + 1]; b_in0 = off_in[b]; b_sz = off_in[b + 1] - b_in0; }
        uint32_t j = o - b_out0, i0 = b_in0 + 2 * j;
        bool has1 = 2 * j + 1 < b_sz;
        Affine<F> p0 = io.load(i0), p1 = has1 ? io.load(i0 + 1) : Affine<F>::inf();
        F den;
        pair_classify(p0, p1, has1, den);
        scratch[o] = run;
        run = run * den;
    }
    F inv = run.inverse();
    // ---- backward: peel the inverses, finish the additions ----
    for (uint32_t o = o1; o-- > o0;) {
        while (o < b_out0) { b--; b_end = b_out0; b_out0 = off_out[b]; b_in0 = off_in[b]; b_sz = off_in[b + 1] - b_in0; }
        uint32_t j = o - b_out0, i0 = b_in0 + 2 * j;
        bool has1 = 2 * j + 1 < b_sz;
        Affine<F> p0 = io.load(i0), p1 = has1 ? io.load(i0 + 1) : Affine<F>::inf();
        F den;
        int mode = pair_classify(p0, p1, has1, den);
        F dinv = inv * scratch[o];
        inv = inv * den;
        out_pts[o] = pair_finish(mode, p0, p1, dinv);
    }
}

}  // namespace zkmsm
