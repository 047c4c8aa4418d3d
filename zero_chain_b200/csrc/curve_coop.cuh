// Warp-cooperative XYZZ group operations for the SERIAL tails of the MSM (Horner over windows, doubling chains of the bucket
// reduction, the prover's variable-base multiplications, the fold of the ranks' partial sums).
//
// On one GPU thread a Montgomery product is a ~480-instruction dependent chain (~1.3 us: every instruction waits for the one
// before), so a doubling (9 products) costs ~12 us and a chain of 240 doublings ~3 ms with 31 lanes of the warp idle.  The
// products of ONE doubling / addition are largely independent of each other, so here every lane of the warp holds the same
// point, each of the first few lanes computes one product of the current stage, and the results are broadcast with shuffles:
// a doubling is 3 dependent stages instead of 9 products, a full addition 4 instead of 14.  Same formulas as XYZZ::dbl /
// XYZZ::add (curve.cuh; dbl-2008-s-1, add-2008-s), same exceptional cases, hence the same group elements.
// All 32 lanes must call these functions together with identical arguments (the control flow is warp-uniform).
#pragma once
#include "curve.cuh"

namespace zkcoop {

template <class F>
__device__ __forceinline__ F bcast(const F &v, int src) {
    F r;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&v);
    uint32_t *d = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 4); k++) d[k] = __shfl_sync(0xffffffffu, s[k], src);
    return r;
}
// operand of this lane: c0 for lane 0, c1 for lane 1, ... (lanes beyond the list compute a throw-away product of c0)
template <class F>
__device__ __forceinline__ F pick(int lane, const F &c0, const F &c1, const F &c2, const F &c3) {
    F r;
    const uint32_t *p0 = reinterpret_cast<const uint32_t *>(&c0), *p1 = reinterpret_cast<const uint32_t *>(&c1);
    const uint32_t *p2 = reinterpret_cast<const uint32_t *>(&c2), *p3 = reinterpret_cast<const uint32_t *>(&c3);
    uint32_t *d = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(F) / 4); k++) d[k] = lane == 1 ? p1[k] : (lane == 2 ? p2[k] : (lane == 3 ? p3[k] : p0[k]));
    return r;
}

// P <- 2 P
template <class F>
__device__ __noinline__ void dbl(XYZZ<F> &P) {
    if (P.is_inf()) return;
    const int lane = threadIdx.x & 31;
    const F U = P.y.dbl();
    // stage A: V = U^2, XX = X^2
    F a = pick(lane, U, P.x, U, U), p = a * a;
    const F V = bcast(p, 0), XX = bcast(p, 1);
    const F M = XX.dbl() + XX;
    // stage B: W = U V, S = X V, MM = M^2
    a = pick(lane, U, P.x, M, U);
    F b = pick(lane, V, V, M, V);
    p = a * b;
    const F W = bcast(p, 0), S = bcast(p, 1), MM = bcast(p, 2);
    const F X3 = MM - S.dbl();
    // stage C: M (S - X3), W Y, V ZZ, W ZZZ
    a = pick(lane, M, W, V, W);
    b = pick(lane, S - X3, P.y, P.zz, P.zzz);
    p = a * b;
    const F t0 = bcast(p, 0), t1 = bcast(p, 1);
    P.zz = bcast(p, 2); P.zzz = bcast(p, 3);
    P.x = X3; P.y = t0 - t1;
}

// P <- P + Q
template <class F>
__device__ __noinline__ void add(XYZZ<F> &P, const XYZZ<F> &Q) {
    if (Q.is_inf()) return;
    if (P.is_inf()) { P = Q; return; }
    const int lane = threadIdx.x & 31;
    // stage 1: U1 = X1 ZZ2, U2 = X2 ZZ1, S1 = Y1 ZZZ2, S2 = Y2 ZZZ1
    F a = pick(lane, P.x, Q.x, P.y, Q.y), b = pick(lane, Q.zz, P.zz, Q.zzz, P.zzz), p = a * b;
    const F U1 = bcast(p, 0), U2 = bcast(p, 1), S1 = bcast(p, 2), S2 = bcast(p, 3);
    const F Pd = U2 - U1, R = S2 - S1;
    if (Pd.is_zero()) {
        if (R.is_zero()) dbl(P); else P = XYZZ<F>::inf();
        return;
    }
    // stage 2: PP = Pd^2, RR = R^2, ZZ12 = ZZ1 ZZ2, ZZZ12 = ZZZ1 ZZZ2
    a = pick(lane, Pd, R, P.zz, P.zzz); b = pick(lane, Pd, R, Q.zz, Q.zzz); p = a * b;
    const F PP = bcast(p, 0), RR = bcast(p, 1), ZZ12 = bcast(p, 2), ZZZ12 = bcast(p, 3);
    // stage 3: PPP = Pd PP, Qv = U1 PP, ZZ3 = ZZ12 PP
    a = pick(lane, Pd, U1, ZZ12, Pd); p = a * PP;
    const F PPP = bcast(p, 0), Qv = bcast(p, 1);
    P.zz = bcast(p, 2);
    const F X3 = RR - PPP - Qv.dbl();
    // stage 4: R (Qv - X3), S1 PPP, ZZZ3 = ZZZ12 PPP
    a = pick(lane, R, S1, ZZZ12, R); b = pick(lane, Qv - X3, PPP, PPP, PPP); p = a * b;
    const F t0 = bcast(p, 0), t1 = bcast(p, 1);
    P.zzz = bcast(p, 2);
    P.x = X3; P.y = t0 - t1;
}

}  // namespace zkcoop
