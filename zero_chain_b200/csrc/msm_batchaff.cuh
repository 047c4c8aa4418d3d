// Batched-affine bucket reduction for the Pippenger MSM (see msm.cuh for where it sits in the schedule).
//
// A mixed XYZZ addition costs 10 Montgomery products.  An AFFINE addition costs one inversion + 3 products
// (lambda = dy / dx, lambda^2, lambda * (x1 - x3)), and Montgomery's simultaneous-inversion trick shares the inversion:
// for denominators d_0 .. d_{m-1} one inverts their product once and peels the individual inverses off with 3 more
// products each — 6 products per addition.  The round-1 experiment put one binary-Euclid inversion in every THREAD
// (64 additions each): the data-dependent Euclid loops diverge inside a warp and the measured cost was 13
// product-equivalents per addition.  Here ONE inversion serves a whole ROUND, and a round is three kernels:
//
//   k_ba_forward   thread: K consecutive output slots; classifies each pair, multiplies the denominators into a running
//                  product, parks the exclusive prefix products (k-major, coalesced) and the pair sources; warp: two shuffle
//                  scans give every thread the product of the OTHER 31 thread totals, and the warp its total
//   k_ba_invert    ONE block over the warp totals: serial chunks + a product tree, a single field inversion (binary
//                  extended Euclid on one thread, ~90 us — the only serial step of the round), and the way back down
//   k_ba_backward  thread: 1 / (own total) = 1 / (warp total) x (product of the others); peels the inverse of each
//                  denominator (2 products), finishes the affine addition (3 products) and stores the sum
//
// A round halves every bucket: bucket b with m points yields ceil(m/2) points (an odd leftover is copied), so the outputs
// are again grouped by bucket and the offsets come from one scan.  After `levels` rounds the (short) remainders go through
// the XYZZ accumulation as before.  Same group elements as bellman's bucket sums, so the canonical result
// cannot change; the exceptional cases the reference's addition handles (P + P -> double, P + (-P) -> infinity, infinity
// operands; ec.rs:357-365, 394-397, 447-456, 473-476) are classified per pair in msm_affine_core.cuh, and the denominator of
// a pair that needs no division is 1, so the shared product is never zero.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "curve.cuh"
#include "msm_affine_core.cuh"
#include "msm_warp_scan.cuh"

namespace zkmsm {

constexpr int BA_T = 128;                     // threads per block of the forward / backward kernels
constexpr int BA_K = 32;                      // additions per thread and round (measured: 32 beats 16 / adaptive once the block trees are gone)
constexpr int BA_MINB = 4;                    // resident blocks per SM of the backward kernel (128 registers; measured 1-3 % faster than 3)
constexpr int BA_MAX_LEVELS = 8;
constexpr int BA_INV_T = 512;                 // threads of the single inversion block
constexpr uint32_t BA_NONE = 0xffffffffu;     // "no second point": the odd leftover of a bucket

// sizes_out[b] = ceil(size_in[b] / 2)
static __global__ void k_half_sizes(const uint32_t *__restrict__ off_in, uint32_t *__restrict__ sizes_out, uint32_t n_buckets) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_buckets) sizes_out[b] = (off_in[b + 1] - off_in[b] + 1) >> 1;
}

template <class F>
__device__ __forceinline__ void ba_store_f(F *dst, const F &v) {
    const uint4 *s = reinterpret_cast<const uint4 *>(&v);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) d[k] = s[k];
}
template <class F>
__device__ __forceinline__ F ba_load_f(const F *src) {
    F v;
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(&v);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) d[k] = s[k];
    return v;
}
// point `code` of the round's input: FIRST round = window-table row (code & 0x7fffffff), negated when bit 31 is set;
// later rounds = position in the previous round's output
template <class F, bool FIRST>
__device__ __forceinline__ const Affine<F> *ba_addr(const Affine<F> *pts, uint32_t code) { return pts + (FIRST ? (code & 0x7fffffffu) : code); }
template <class F, bool FIRST>
__device__ __forceinline__ Affine<F> ba_load_point(const Affine<F> *pts, uint32_t code) {
    Affine<F> p;
    const uint4 *s = reinterpret_cast<const uint4 *>(ba_addr<F, FIRST>(pts, code));
    uint4 *d = reinterpret_cast<uint4 *>(&p);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(Affine<F>) / 16); k++) d[k] = __ldg(s + k);
    if (FIRST) p.y = p.y.cneg(code >> 31);
    return p;
}

// ---- forward ------------------------------------------------------------------------------------------------------
// in_pts: FIRST ? window tables : previous round's points.  sorted: FIRST only (entry codes grouped by bucket).
// prefix: [(K + 1)][T_total] elements (k-major; plane K holds the thread totals), srcs: [n_out] pair sources.
// The denominator of an ordinary pair is x1 - x0, so this pass gathers only the x coordinates (half the bytes); the rare pairs
// that need more (equal x: doubling or cancellation; x = 0: possibly the point at infinity) fetch the full points.
template <class F, bool FIRST>
__device__ __forceinline__ F ba_load_x(const Affine<F> *pts, uint32_t code) {
    F x;
    const uint4 *s = reinterpret_cast<const uint4 *>(ba_addr<F, FIRST>(pts, code));
    uint4 *d = reinterpret_cast<uint4 *>(&x);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 16); k++) d[k] = __ldg(s + k);
    return x;
}
template <class F, bool FIRST>
__global__ void __launch_bounds__(BA_T, 4) k_ba_forward(const Affine<F> *__restrict__ in_pts, const uint32_t *__restrict__ sorted,
                                                         const uint32_t *__restrict__ off_in, const uint32_t *__restrict__ off_out, uint32_t n_buckets, int K,
                                                         F *__restrict__ prefix, uint2 *__restrict__ srcs, F *__restrict__ block_totals) {
    const uint32_t total = off_out[n_buckets];
    const uint32_t n_blocks = (total + BA_T * K - 1) / (BA_T * K);
    if (blockIdx.x >= n_blocks) return;                    // the grid is sized for the host-side upper bound of `total`
    const int t = threadIdx.x;
    const size_t T_total = (size_t)gridDim.x * BA_T, tid = (size_t)blockIdx.x * BA_T + t;
    const uint32_t o0 = (uint32_t)tid * K;
    F run = F::one();
    if (o0 < total) {
        const uint32_t o1 = o0 + K < total ? o0 + K : total;
        {   // pass A: the sources of this thread's output slots (walk over the buckets; entry codes read in order)
            uint32_t lo = 0, hi = n_buckets;               // bucket of o0: last b with off_out[b] <= o0
            while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (off_out[mid] <= o0) lo = mid; else hi = mid; }
            uint32_t b = lo, b_end = off_out[b + 1], b_out0 = off_out[b], b_in0 = off_in[b], b_sz = off_in[b + 1] - b_in0;
            for (uint32_t o = o0; o < o1; o++) {
                while (o >= b_end) { b++; b_out0 = b_end; b_end = off_out[b + 1]; b_in0 = off_in[b]; b_sz = off_in[b + 1] - b_in0; }
                const uint32_t j = o - b_out0, i0 = b_in0 + 2 * j;
                const bool has1 = 2 * j + 1 < b_sz;
                uint2 src;
                src.x = FIRST ? sorted[i0] : i0;
                src.y = has1 ? (FIRST ? sorted[i0 + 1] : i0 + 1) : BA_NONE;
                srcs[o] = src;
            }
        }
        // pass B: denominators and their running product, the x gathers issued two additions ahead
        uint2 s0 = srcs[o0], s1 = o0 + 1 < o1 ? srcs[o0 + 1] : make_uint2(0, BA_NONE);
        F xa0 = ba_load_x<F, FIRST>(in_pts, s0.x), xa1 = s0.y != BA_NONE ? ba_load_x<F, FIRST>(in_pts, s0.y) : F::zero();
        F xb0 = F::zero(), xb1 = F::zero();
        if (o0 + 1 < o1) { xb0 = ba_load_x<F, FIRST>(in_pts, s1.x); if (s1.y != BA_NONE) xb1 = ba_load_x<F, FIRST>(in_pts, s1.y); }
        for (uint32_t o = o0; o < o1; o++) {
            const uint2 cur = s0;
            const F x0 = xa0, x1 = xa1;
            s0 = s1; xa0 = xb0; xa1 = xb1;
            if (o + 2 < o1) {
                s1 = srcs[o + 2];
                xb0 = ba_load_x<F, FIRST>(in_pts, s1.x);
                if (s1.y != BA_NONE) xb1 = ba_load_x<F, FIRST>(in_pts, s1.y);
            }
            ba_store_f(prefix + (size_t)(o - o0) * T_total + tid, run);
            if (cur.y == BA_NONE) continue;                 // odd leftover: copied by the backward pass
            F den = x1 - x0;
            if (den.is_zero() || x0.is_zero() || x1.is_zero()) {       // rare: decide on the full points, exactly as the backward pass will
                Affine<F> p0 = ba_load_point<F, FIRST>(in_pts, cur.x), p1 = ba_load_point<F, FIRST>(in_pts, cur.y);
                if (pair_classify(p0, p1, true, den) > PAIR_DBL) continue;
            }
            run = run * den;
        }
    }
    // the product of the OTHER 31 thread totals of this warp goes to plane K; the warp's total to the round's inversion
    F others, all;
    ba_warp_products(run, others, all);
    ba_store_f(prefix + (size_t)K * T_total + tid, others);
    if ((t & 31) == 0) ba_store_f(block_totals + (tid >> 5), all);
}

// ---- the round's single inversion ----------------------------------------------------------------------------------
// inv_out[i] = 1 / totals[i] for i < n = ceil(off_out[n_buckets] / (BA_T K)); scratch: n elements.
template <class F>
__global__ void __launch_bounds__(BA_INV_T) k_ba_invert(const F *__restrict__ totals, const uint32_t *__restrict__ off_out, uint32_t n_buckets, int K,
                                                        F *__restrict__ scratch, F *__restrict__ inv_out) {
    extern __shared__ unsigned char ba_smem[];
    F *node = reinterpret_cast<F *>(ba_smem);                // heap of 2 * BA_INV_T nodes
    const uint32_t total = off_out[n_buckets];
    const uint32_t n = ((total + BA_T * K - 1) / (BA_T * K)) * (BA_T / 32);       // one total per warp of the live blocks
    const int t = threadIdx.x;
    const uint32_t per = (n + BA_INV_T - 1) / BA_INV_T, c0 = t * per, c1 = c0 + per < n ? c0 + per : n;
    F run = F::one();
    for (uint32_t i = c0; i < c1; i++) { ba_store_f(scratch + i, run); run = run * ba_load_f(totals + i); }
    node[BA_INV_T + t] = run;
    for (int s = BA_INV_T / 2; s >= 1; s >>= 1) {
        __syncthreads();
        if (t < s) node[s + t] = node[2 * (s + t)] * node[2 * (s + t) + 1];
    }
    __syncthreads();
    if (t == 0) node[1] = node[1].inverse();                  // every factor is non-zero by construction (pair_classify)
    for (int s = 1; s < BA_INV_T; s <<= 1) {
        __syncthreads();
        if (t < s) {
            F inv = node[s + t], l = node[2 * (s + t)], r = node[2 * (s + t) + 1];
            node[2 * (s + t)] = inv * r;
            node[2 * (s + t) + 1] = inv * l;
        }
    }
    __syncthreads();
    F inv = node[BA_INV_T + t];
    for (uint32_t i = c1; i-- > c0;) {
        F v = ba_load_f(totals + i);
        ba_store_f(inv_out + i, inv * ba_load_f(scratch + i));
        inv = inv * v;
    }
}

// ---- backward ------------------------------------------------------------------------------------------------------
template <class F, bool FIRST, int MINB>
__global__ void __launch_bounds__(BA_T, MINB) k_ba_backward(const Affine<F> *__restrict__ in_pts, const uint32_t *__restrict__ off_out, uint32_t n_buckets, int K,
                                                             const F *__restrict__ prefix, const uint2 *__restrict__ srcs, const F *__restrict__ block_inv,
                                                             Affine<F> *__restrict__ out_pts) {
    constexpr int PV = (int)(sizeof(Affine<F>) / 16), FV = (int)(sizeof(F) / 16);     // a slot = two points + one prefix element
    extern __shared__ unsigned char ba_smem[];
    uint4 *stage = reinterpret_cast<uint4 *>(ba_smem);
    const uint32_t total = off_out[n_buckets];
    const uint32_t n_blocks = (total + BA_T * K - 1) / (BA_T * K);
    if (blockIdx.x >= n_blocks) return;
    const int t = threadIdx.x;
    const size_t T_total = (size_t)gridDim.x * BA_T, tid = (size_t)blockIdx.x * BA_T + t;
    const uint32_t o0 = (uint32_t)tid * K;
    if (o0 >= total) return;
    // inverse of this thread's total = (inverse of the warp's total) x (product of the other 31 totals of the warp)
    F inv = ba_load_f(block_inv + (tid >> 5)) * ba_load_f(prefix + (size_t)K * T_total + tid);
    const uint32_t o1 = o0 + K < total ? o0 + K : total;
    // software pipeline: while addition o is finished, the operands of o - 1 (two points, one prefix product) stream into
    // this thread's shared-memory slot with cp.async — the gather latency hides behind five Montgomery products
    uint4 *slot = stage + t;                                  // vector v of the slot lives at stage[v * BA_T + t]: conflict-free
    auto prefetch = [&](uint32_t o) {
        uint2 src = srcs[o];
        const uint4 *a = reinterpret_cast<const uint4 *>(ba_addr<F, FIRST>(in_pts, src.x));
        const uint4 *b = reinterpret_cast<const uint4 *>(ba_addr<F, FIRST>(in_pts, src.y == BA_NONE ? src.x : src.y));
        const uint4 *c = reinterpret_cast<const uint4 *>(prefix + (size_t)(o - o0) * T_total + tid);
#pragma unroll
        for (int v = 0; v < PV; v++) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(slot + v * BA_T)), "l"(a + v) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(slot + (PV + v) * BA_T)), "l"(b + v) : "memory");
        }
#pragma unroll
        for (int v = 0; v < FV; v++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(slot + (2 * PV + v) * BA_T)), "l"(c + v) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        return src;
    };
    uint2 src = prefetch(o1 - 1);
    for (uint32_t o = o1; o-- > o0;) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        Affine<F> p0, p1;
        F pre;
        {
            uint4 *d0 = reinterpret_cast<uint4 *>(&p0), *d1 = reinterpret_cast<uint4 *>(&p1), *d2 = reinterpret_cast<uint4 *>(&pre);
#pragma unroll
            for (int v = 0; v < PV; v++) { d0[v] = slot[v * BA_T]; d1[v] = slot[(PV + v) * BA_T]; }
#pragma unroll
            for (int v = 0; v < FV; v++) d2[v] = slot[(2 * PV + v) * BA_T];
        }
        const uint2 cur = src;
        if (o > o0) src = prefetch(o - 1);
        const bool has1 = cur.y != BA_NONE;
        if (FIRST) { p0.y = p0.y.cneg(cur.x >> 31); if (has1) p1.y = p1.y.cneg(cur.y >> 31); }
        if (!has1) p1 = Affine<F>::inf();
        F den;
        const int mode = pair_classify(p0, p1, has1, den);
        F dinv = F::one();
        if (mode <= PAIR_DBL) { dinv = inv * pre; inv = inv * den; }
        Affine<F> r = pair_finish(mode, p0, p1, dinv);
        uint4 *dst = reinterpret_cast<uint4 *>(out_pts + o);
        const uint4 *rs = reinterpret_cast<const uint4 *>(&r);
#pragma unroll
        for (int v = 0; v < PV; v++) dst[v] = rs[v];
    }
}

// dynamic shared memory of the three kernels
template <class F> constexpr size_t ba_smem_forward() { return 0; }
template <class F> constexpr size_t ba_smem_backward() { return (size_t)BA_T * (2 * sizeof(Affine<F>) + sizeof(F)); }
template <class F> constexpr size_t ba_smem_invert() { return 2 * BA_INV_T * sizeof(F); }

}  // namespace zkmsm
