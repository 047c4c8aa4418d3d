// Per-pair pieces of the batched-affine bucket reduction (see msm_batchaff.cuh): classification of a pair,
// its denominator, and the affine addition / doubling given the inverted denominator.  Plain device functions
// with no CUDA-runtime dependency so the host emulation (tests/host_emul) compiles the same source.
#pragma once
#include "curve.cuh"

namespace zkmsm {

enum { PAIR_ADD = 0, PAIR_DBL = 1, PAIR_COPY0 = 2, PAIR_COPY1 = 3, PAIR_INF = 4 };

template <class F>
struct PairIO {
    const Affine<F> *pts;          // level input points (the window tables at level 0)
    const uint32_t *sorted;        // level 0: entry codes (table index | sign << 31); nullptr above
    ZK_DEV Affine<F> load(uint32_t i) const {
        Affine<F> p;
        if (sorted) {
            uint32_t code = sorted[i];
            p = pts[code & 0x7fffffffu];
            p.y = p.y.cneg(code >> 31);
        } else p = pts[i];
        return p;
    }
};

// classify the pair (p0, p1 or nothing) and return the denominator whose inverse the addition needs
template <class F>
ZK_DEV int pair_classify(const Affine<F> &p0, const Affine<F> &p1, bool has1, F &den) {
    den = F::one();
    if (!has1 || p1.is_inf()) return PAIR_COPY0;
    if (p0.is_inf()) return PAIR_COPY1;
    if (p0.x == p1.x) {
        if (p0.y == p1.y) {
            den = p0.y.dbl();
            if (!den.is_zero()) return PAIR_DBL;
            den = F::one();                                             // y = 0: a 2-torsion point (only an unchecked, malformed base can be one): 2P = O,
            return PAIR_INF;                                            // and the shared product of denominators must stay non-zero
        }
        return PAIR_INF;                                                // P + (-P)
    }
    den = p1.x - p0.x;
    return PAIR_ADD;
}
template <class F>
ZK_DEV Affine<F> pair_finish(int mode, const Affine<F> &p0, const Affine<F> &p1, const F &dinv) {
    if (mode == PAIR_COPY0) return p0;
    if (mode == PAIR_COPY1) return p1;
    if (mode == PAIR_INF) return Affine<F>::inf();
    F lam;
    if (mode == PAIR_DBL) { F xx = p0.x.sqr(); lam = (xx.dbl() + xx) * dinv; }
    else lam = (p1.y - p0.y) * dinv;
    Affine<F> r;
    r.x = lam.sqr() - p0.x - p1.x;          // DBL: p1.x == p0.x
    r.y = lam * (p0.x - r.x) - p0.y;
    return r;
}

}  // namespace zkmsm
