// Lane-parallel Fq12 arithmetic for the Groth16 verifier's long loops.
//
// The thread-per-proof Miller loop / final exponentiation of pairing.cuh keep an Fq12 value (144 registers) plus its
// temporaries per thread: ptxas spills them (thousands of LDL/STL, 18 GB of local-memory traffic per 32 k proofs) and a
// block-sized batch is bound by the latency of one thread's chain.  Here an Fq12 value is spread over SIX lanes of a warp:
//   f = sum_t c_t w^t,  w^6 = xi = 1 + u,  c_t in Fq2        (slot t = 2 j + i of the tower element (w^i, v^j), pairing.cuh)
// lane t holds c_t (24 registers).  Products are the cyclic convolution c_t = sum_{i+j=t} a_i b_j + xi sum_{i+j=t+6} a_i b_j with
// operands fetched by warp shuffles, so nothing lives in local memory and six lanes share the serial chain:
//   general product    6 Fq2 products per lane   (thread version: 18, one after the other)
//   squaring           4 (symmetric pairs once)  (12)
//   sparse line "014"  3                         (13)
//   cyclotomic square  2 Fq2 squarings           (9)
// A warp carries five such groups (lanes 30, 31 shadow group 4).  The values are the same field elements as in the
// thread version (same tower, same formulas up to the order of additions, all results fully reduced), so verdicts and
// Fq12 bytes are identical — tests/test_gpu_verify.py compares both with the oracle.
#pragma once
#include "pairing.cuh"

namespace zklanes {
using zkpair::mul_xi;
using zkpair::conj2;

constexpr int GROUPS_PER_WARP = 5;

struct Lane {
    int base;      // first lane of this group inside the warp
    int t;         // slot 0..5
    bool live;     // lanes 30, 31 mirror lanes 24, 25 and never store
    __device__ __forceinline__ static Lane make() {
        const int lane = threadIdx.x & 31;
        Lane L;
        int g = lane / 6;
        L.live = g < GROUPS_PER_WARP;
        if (!L.live) g = GROUPS_PER_WARP - 1;
        L.t = L.live ? lane - 6 * g : lane - 30;
        L.base = 6 * g;
        return L;
    }
};

__device__ __forceinline__ Fq shfl_fq(const Fq &v, int src) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = __shfl_sync(0xffffffffu, v.l[k], src);
    return r;
}
__device__ __forceinline__ Fq2 shfl2(const Fq2 &v, int src) { Fq2 r; r.c0 = shfl_fq(v.c0, src); r.c1 = shfl_fq(v.c1, src); return r; }
__device__ __forceinline__ Fq2 sel2(bool c, const Fq2 &a, const Fq2 &b) {
    Fq2 r;
#pragma unroll
    for (int k = 0; k < 12; k++) { r.c0.l[k] = c ? a.c0.l[k] : b.c0.l[k]; r.c1.l[k] = c ? a.c1.l[k] : b.c1.l[k]; }
    return r;
}

// c_t of a * b
__device__ __noinline__ Fq2 mul12(const Lane L, const Fq2 a, const Fq2 b) {      // by value: operands travel in registers, one copy of the code
    Fq2 acc = Fq2::zero();
#pragma unroll 1
    for (int i = 0; i < 6; i++) {
        int j = L.t - i;
        const bool wrap = j < 0;
        if (wrap) j += 6;
        Fq2 p = shfl2(a, L.base + i) * shfl2(b, L.base + j);
        acc = acc + sel2(wrap, mul_xi(p), p);
    }
    return acc;
}
// c_t of a^2: unordered pairs {i, j}, i + j = t (mod 6); byte = i | j << 3 | (i < j) << 6 | (i + j >= 6) << 7, 0xff = no fourth pair
static __device__ __constant__ uint8_t ZK_SQR_PAIRS[6][4] = {
    {0 | 0 << 3, 1 | 5 << 3 | 0xc0, 2 | 4 << 3 | 0xc0, 3 | 3 << 3 | 0x80},
    {0 | 1 << 3 | 0x40, 2 | 5 << 3 | 0xc0, 3 | 4 << 3 | 0xc0, 0xff},
    {0 | 2 << 3 | 0x40, 1 | 1 << 3, 3 | 5 << 3 | 0xc0, 4 | 4 << 3 | 0x80},
    {0 | 3 << 3 | 0x40, 1 | 2 << 3 | 0x40, 4 | 5 << 3 | 0xc0, 0xff},
    {0 | 4 << 3 | 0x40, 1 | 3 << 3 | 0x40, 2 | 2 << 3, 5 | 5 << 3 | 0x80},
    {0 | 5 << 3 | 0x40, 1 | 4 << 3 | 0x40, 2 | 3 << 3 | 0x40, 0xff}};
__device__ __forceinline__ Fq2 sqr12(const Lane &L, const Fq2 &a) {
    Fq2 acc = Fq2::zero();
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        const uint8_t e = ZK_SQR_PAIRS[L.t][s];
        const bool none = e == 0xff;
        const int i = none ? 0 : (e & 7), j = none ? 0 : ((e >> 3) & 7);
        Fq2 p = shfl2(a, L.base + i) * shfl2(a, L.base + j);
        p = sel2((e & 0x40) != 0, p.dbl(), p);           // per-lane choices are selects; the product above is warp-uniform code
        p = sel2((e & 0x80) != 0, mul_xi(p), p);
        acc = acc + sel2(none, Fq2::zero(), p);
    }
    return acc;
}
// c_t of f * (s0 + s1 w^2 + s4 w^3)      (the line through T evaluated at P: slots 0, 2, 3 — "014" in tower numbering)
__device__ __forceinline__ Fq2 mul12_line(const Lane &L, const Fq2 &f, const Fq2 &s0, const Fq2 &s1, const Fq2 &s4) {
    Fq2 acc = f * s0;
    {
        int j = L.t - 2; const bool wrap = j < 0; if (wrap) j += 6;
        Fq2 p = shfl2(f, L.base + j) * s1;
        acc = acc + sel2(wrap, mul_xi(p), p);
    }
    {
        int j = L.t - 3; const bool wrap = j < 0; if (wrap) j += 6;
        Fq2 p = shfl2(f, L.base + j) * s4;
        acc = acc + sel2(wrap, mul_xi(p), p);
    }
    return acc;
}
// f * line with the G1 point folded in: s0 = c2, s1 = c1 * xP, s4 = c0 * yP.  The four Fq products of the two scalings are
// spread over lanes 0..3 of the group and broadcast; a lane loads only the coefficient words it needs (its scaling operand and c2).
// P = (xP, yP) sits in shared memory (pxy[0], pxy[1]): six lanes share one point and the registers are needed elsewhere.
__device__ __forceinline__ Fq ldg_fq(const Fq *p) {
    Fq r;
    const uint4 *s = reinterpret_cast<const uint4 *>(p);
    uint4 *d = reinterpret_cast<uint4 *>(&r);
#pragma unroll
    for (int k = 0; k < 3; k++) d[k] = __ldg(s + k);
    return r;
}
__device__ __forceinline__ Fq2 ell(const Lane &L, const Fq2 &f, const zkpair::LineCoeff *c, const Fq *pxy) {
    const Fq *w = reinterpret_cast<const Fq *>(c);         // c0.c0, c0.c1, c1.c0, c1.c1, c2.c0, c2.c1
    const int idx = L.t == 0 ? 2 : (L.t == 1 ? 3 : (L.t == 2 ? 0 : 1));
    const Fq p = ldg_fq(w + idx) * pxy[L.t < 2 ? 0 : 1];
    Fq2 s1, s4, s0;
    s1.c0 = shfl_fq(p, L.base + 0); s1.c1 = shfl_fq(p, L.base + 1);
    s4.c0 = shfl_fq(p, L.base + 2); s4.c1 = shfl_fq(p, L.base + 3);
    s0.c0 = ldg_fq(w + 4); s0.c1 = ldg_fq(w + 5);
    return mul12_line(L, f, s0, s1, s4);
}
// The Miller loops of the three pairs of one proof with ONE accumulator (what Engine::miller_loop does for a slice of pairs,
// mod.rs:60-113): per bit the three lines are multiplied in and f is squared once — the thread version runs three separate loops
// (three squarings per bit) to have 3 n work items.  A pair with a point at infinity contributes 1: its lines are skipped.
struct PairIn { const Fq *pxy; const zkpair::LineCoeff *coeffs; size_t stride; bool skip; };
__device__ __forceinline__ Fq2 ell_or_skip(const Lane &L, const Fq2 &f, const PairIn &p, int n) {
    Fq2 g = ell(L, f, p.coeffs + (size_t)n * p.stride, p.pxy);
    return sel2(p.skip, f, g);
}
__device__ __forceinline__ Fq2 miller_loop3(const Lane &L, const PairIn &p0, const PairIn &p1, const PairIn &p2) {
    Fq2 f = L.t == 0 ? Fq2::one() : Fq2::zero();
    int n = 0;
#pragma unroll 1
    for (int i = 62; i >= 1; i--) {
        f = ell_or_skip(L, f, p0, n); f = ell_or_skip(L, f, p1, n); f = ell_or_skip(L, f, p2, n); n++;
        if ((zkpair::BLS_X_ABS >> i) & 1) { f = ell_or_skip(L, f, p0, n); f = ell_or_skip(L, f, p1, n); f = ell_or_skip(L, f, p2, n); n++; }
        f = sqr12(L, f);
    }
    f = ell_or_skip(L, f, p0, n); f = ell_or_skip(L, f, p1, n); f = ell_or_skip(L, f, p2, n);
    return (L.t & 1) ? f.neg() : f;                       // conj: w -> -w
}

// ---- final exponentiation (pairing.cuh final_exponentiation, same chain) ---------------------------------------------
__device__ __forceinline__ Fq2 conj12(const Lane &L, const Fq2 &f) { return (L.t & 1) ? f.neg() : f; }
// x * w^2 (= multiplication by v): slot t <- slot t - 2, times xi on wrap-around
__device__ __forceinline__ Fq2 mul_w2(const Lane &L, const Fq2 &x) {
    int j = L.t - 2; const bool wrap = j < 0; if (wrap) j += 6;
    Fq2 p = shfl2(x, L.base + j);
    return sel2(wrap, mul_xi(p), p);
}
__device__ __forceinline__ bool all_lanes(const Lane &L, bool v) {      // AND over the six lanes of the group
    const unsigned m = __ballot_sync(0xffffffffu, v);
    return ((m >> L.base) & 0x3fu) == 0x3fu;
}
// gamma_k^t for this lane's slot, k = 1..3 (pairing.cuh frobenius12)
struct FrobConsts { Fq2 g[3]; };
__device__ __forceinline__ FrobConsts frob_consts(const Lane &L) {
    FrobConsts c;
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        const Fq2 g = zkpair::frob_gamma(k + 1);
        Fq2 pw = Fq2::one();
#pragma unroll 1
        for (int e = 1; e <= 5; e++) pw = sel2(e <= L.t, pw * g, pw);
        c.g[k] = pw;
    }
    return c;
}
__device__ __forceinline__ Fq2 frobenius12(const Lane &L, const FrobConsts &fc, const Fq2 &f, int k) {
    Fq2 c = (k & 1) ? conj2(f) : f;
    Fq2 r = c * fc.g[k - 1];
    return sel2(L.t == 0, c, r);
}
// 1 / f through the quadratic tower f = a0 + a1 w (a0: even slots, a1: odd slots, both in Fq6 = Fq2[v], v = w^2):
// 1 / f = conj(f) / (a0^2 - v a1^2); the Fq6 inverse of the norm by the adjugate formulas, every lane redundantly
__device__ __forceinline__ Fq2 inv12(const Lane &L, const Fq2 &f) {
    const bool even = (L.t & 1) == 0;
    const Fq2 zero = Fq2::zero();
    const Fq2 a0 = sel2(even, f, zero);                                   // a0 embedded at the even slots
    const Fq2 up = shfl2(f, L.base + (L.t < 5 ? L.t + 1 : 5));            // slot t + 1
    const Fq2 a1 = sel2(even, up, zero);                                  // a1 shifted down onto the even slots
    const Fq2 nrm = sqr12(L, a0) - mul_w2(L, sqr12(L, a1));               // a0^2 - v a1^2 (even slots)
    const Fq2 n0 = shfl2(nrm, L.base + 0), n1 = shfl2(nrm, L.base + 2), n2 = shfl2(nrm, L.base + 4);
    const Fq2 A = n0.sqr() - mul_xi(n1 * n2), B = mul_xi(n2.sqr()) - n0 * n1, C = n1.sqr() - n0 * n2;
    const Fq2 d = (n0 * A + mul_xi(n2 * B + n1 * C)).inverse();
    const Fq2 mine = L.t == 0 ? A : (L.t == 2 ? B : C);
    const Fq2 ninv = sel2(even, mine * d, zero);                          // 1 / norm embedded at the even slots
    return mul12(L, conj12(L, f), ninv);
}
// Granger-Scott squaring in the cyclotomic subgroup (pairing.cuh cyclotomic_sqr): Fq4 pairs (c_0, c_3), (c_1, c_4), (c_2, c_5);
// every lane squares its own slot and the sum of its pair, so the nine Fq2 squarings take two steps
__device__ __noinline__ Fq2 cyclotomic_sqr(const Lane L, const Fq2 g) {
    const bool low = L.t < 3;
    const int partner = L.base + (low ? L.t + 3 : L.t - 3);
    const Fq2 other = shfl2(g, partner);
    const Fq2 own2 = g.sqr(), cross = (g + other).sqr();
    const Fq2 oth2 = shfl2(own2, partner);
    const Fq2 aa = sel2(low, own2, oth2), bb = sel2(low, oth2, own2);
    const Fq2 val = sel2(low, aa + mul_xi(bb), cross - aa - bb);          // low lane: r0 of its pair, high lane: r1
    // pair (0,3) stays; pair (1,4) feeds slots 2, 5; pair (2,5) feeds slots 4, 1
    const int src = L.t == 0 ? 0 : (L.t == 3 ? 3 : (L.t == 2 ? 1 : (L.t == 5 ? 4 : (L.t == 4 ? 2 : 5))));
    Fq2 v = shfl2(val, L.base + src);
    v = sel2(L.t == 1, mul_xi(v), v);
    const Fq2 minus = (v - g).dbl() + v, plus = (v + g).dbl() + v;        // 3 v -/+ 2 c_t
    return sel2((L.t & 1) == 0, minus, plus);
}
__device__ __noinline__ Fq2 exp_x(const Lane L, const Fq2 f) {
    Fq2 r = f;
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
        r = cyclotomic_sqr(L, r);
        if ((zkpair::BLS_X_ABS >> i) & 1) r = mul12(L, r, f);
    }
    return conj12(L, r);
}
// f must be non-zero (checked by the caller)
__device__ __forceinline__ Fq2 final_exponentiation(const Lane &L, const Fq2 &f) {
    const FrobConsts fc = frob_consts(L);
    Fq2 g = mul12(L, conj12(L, f), inv12(L, f));                          // f^(q^6 - 1)
    g = mul12(L, frobenius12(L, fc, g, 2), g);                             // ^(q^2 + 1): cyclotomic subgroup, inverse = conj
    Fq2 a = mul12(L, exp_x(L, g), conj12(L, g));                           // g^(x-1)
    a = mul12(L, exp_x(L, a), conj12(L, a));                               // g^((x-1)^2) = g^l3
    const Fq2 b = exp_x(L, a);                                             // g^l2
    const Fq2 c = mul12(L, exp_x(L, b), conj12(L, a));                     // g^l1
    const Fq2 d = mul12(L, exp_x(L, c), mul12(L, cyclotomic_sqr(L, g), g));   // g^l0
    return mul12(L, mul12(L, d, frobenius12(L, fc, c, 1)), mul12(L, frobenius12(L, fc, b, 2), frobenius12(L, fc, a, 3)));
}

// address of slot t inside a tower-ordered Fq12 (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2)
__device__ __forceinline__ int slot_index(int t) { return (t & 1) * 3 + (t >> 1); }

}  // namespace zklanes
