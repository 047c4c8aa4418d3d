// BLS12-381 pairing for the Groth16 verifier: the Fq12 tower, the prepared form of a G2 point
// (line coefficients), the Miller loop over prepared pairs and the final exponentiation.
//
// What it replaces in the reference: Engine::miller_loop / final_exponentiation / G2Prepared::from_affine
// (core/pairing/src/bls12_381/mod.rs:40-160, 163-359) and the tower arithmetic of fq6.rs / fq12.rs.  Written from the
// mathematics, not from that code:
//   tower      Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), xi = 1 + u   (so w^6 = xi)
//   Frobenius  an element is sum_t c_t w^t (c_t in Fq2, t = 2j + i for the v^j w^i slot); w^(q^k) = gamma_k w with
//              gamma_k = xi^((q^k-1)/6), hence phi^k(sum c_t w^t) = sum conj^k(c_t) gamma_k^t w^t — three constants
//              (pairing_consts.inc) instead of the reference's coefficient tables
//   lines      for T = (X, Y, Z) Jacobian on the twist, untwisting by (x, y) -> (x / w^2, y / w^3) and clearing
//              denominators puts the tangent / chord through T evaluated at P = (xP, yP) in G1 at
//              c2 * 1 + (c1 xP) * v + (c0 yP) * v w      (three of the twelve Fq-pairs: a "014" sparse element)
//              with (c0, c1, c2) exactly the triples G2Prepared stores (scaling included, so prepared keys are
//              byte-identical to PreparedVerifyingKey::write — see pyref.g2_prepare for the closed forms)
//   final exp  easy part f^((q^6-1)(q^2+1)); hard part as 3 (q^4-q^2+1)/r = l0 + l1 q + l2 q^2 + l3 q^3 with
//              l3 = (x-1)^2, l2 = l3 x, l1 = l2 x - l3, l0 = l1 x + 3 (x = -0xd201000000010000): five
//              exponentiations by |x| (Granger-Scott cyclotomic squarings) and three Frobenius maps.  This is the same function of f as the
//              reference's chain (its result is the cube of the plain reduced pairing; fixture conf_vk.dat[0:576]).
// Values are canonical Montgomery field elements throughout, so results are bit-identical to the reference's.
#pragma once
#include "curve.cuh"
#include <stddef.h>
#include "pairing_consts.inc"

#ifdef ZK_HOST_EMUL
static const uint32_t ZK_FROB_GAMMA[3][24] = ZK_FROB_GAMMA_INIT;
#else
static __device__ __constant__ uint32_t ZK_FROB_GAMMA[3][24] = ZK_FROB_GAMMA_INIT;
#endif

namespace zkpair {

constexpr uint64_t BLS_X_ABS = 0xd201000000010000ull;   // mod.rs:23-25; the parameter is -BLS_X_ABS
constexpr int N_COEFFS = 68;                            // 63 doublings + 5 additions (bits set below the top one)

ZK_DEV Fq2 mul_xi(const Fq2 &a) { Fq2 r; r.c0 = a.c0 - a.c1; r.c1 = a.c0 + a.c1; return r; }   // (1+u)(a0 + a1 u)
ZK_DEV Fq2 conj2(const Fq2 &a) { Fq2 r; r.c0 = a.c0; r.c1 = a.c1.neg(); return r; }
ZK_DEV Fq2 mul_fq(const Fq2 &a, const Fq &k) { Fq2 r; r.c0 = a.c0 * k; r.c1 = a.c1 * k; return r; }
ZK_DEV Fq2 frob_gamma(int k) {   // k = 1..3
    Fq2 g;
    for (int i = 0; i < 12; i++) { g.c0.l[i] = ZK_FROB_GAMMA[k - 1][i]; g.c1.l[i] = ZK_FROB_GAMMA[k - 1][12 + i]; }
    return g;
}

struct Fq6 {
    Fq2 c0, c1, c2;   // c0 + c1 v + c2 v^2
    ZK_DEV static Fq6 zero() { Fq6 r; r.c0 = Fq2::zero(); r.c1 = Fq2::zero(); r.c2 = Fq2::zero(); return r; }
    ZK_DEV static Fq6 one() { Fq6 r; r.c0 = Fq2::one(); r.c1 = Fq2::zero(); r.c2 = Fq2::zero(); return r; }
    ZK_DEV bool operator==(const Fq6 &b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
    ZK_DEV bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
    ZK_DEV friend Fq6 operator+(const Fq6 &a, const Fq6 &b) { Fq6 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; r.c2 = a.c2 + b.c2; return r; }
    ZK_DEV friend Fq6 operator-(const Fq6 &a, const Fq6 &b) { Fq6 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; r.c2 = a.c2 - b.c2; return r; }
    ZK_DEV Fq6 neg() const { Fq6 r; r.c0 = c0.neg(); r.c1 = c1.neg(); r.c2 = c2.neg(); return r; }
    ZK_DEV Fq6 mul_v() const { Fq6 r; r.c0 = mul_xi(c2); r.c1 = c0; r.c2 = c1; return r; }   // v^3 = xi
};
// Karatsuba over the three Fq2 slots: 6 Fq2 products
static ZK_PTFN Fq6 mul6(const Fq6 &a, const Fq6 &b) {
    Fq2 t0 = a.c0 * b.c0, t1 = a.c1 * b.c1, t2 = a.c2 * b.c2;
    Fq6 r;
    r.c0 = t0 + mul_xi((a.c1 + a.c2) * (b.c1 + b.c2) - t1 - t2);
    r.c1 = (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1 + mul_xi(t2);
    r.c2 = (a.c0 + a.c2) * (b.c0 + b.c2) - t0 - t2 + t1;
    return r;
}
// a * (b0 + b1 v): 5 Fq2 products
static ZK_PTFN Fq6 mul6_01(const Fq6 &a, const Fq2 &b0, const Fq2 &b1) {
    Fq2 t0 = a.c0 * b0, t1 = a.c1 * b1;
    Fq6 r;
    r.c0 = t0 + mul_xi(a.c2 * b1);
    r.c1 = (a.c0 + a.c1) * (b0 + b1) - t0 - t1;
    r.c2 = a.c2 * b0 + t1;
    return r;
}
// a * (b1 v): 3 Fq2 products
static ZK_PTFN Fq6 mul6_1(const Fq6 &a, const Fq2 &b1) {
    Fq6 r; r.c0 = mul_xi(a.c2 * b1); r.c1 = a.c0 * b1; r.c2 = a.c1 * b1; return r;
}
static ZK_PTFN Fq6 inv6(const Fq6 &a) {
    // adjugate over Fq2: (A, B, C) / (a0 A + xi (a2 B + a1 C))
    Fq2 A = a.c0.sqr() - mul_xi(a.c1 * a.c2);
    Fq2 B = mul_xi(a.c2.sqr()) - a.c0 * a.c1;
    Fq2 C = a.c1.sqr() - a.c0 * a.c2;
    Fq2 n = (a.c0 * A + mul_xi(a.c2 * B + a.c1 * C)).inverse();
    Fq6 r; r.c0 = A * n; r.c1 = B * n; r.c2 = C * n; return r;
}

struct Fq12 {
    Fq6 c0, c1;   // c0 + c1 w
    ZK_DEV static Fq12 one() { Fq12 r; r.c0 = Fq6::one(); r.c1 = Fq6::zero(); return r; }
    ZK_DEV bool operator==(const Fq12 &b) const { return c0 == b.c0 && c1 == b.c1; }
    ZK_DEV bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    ZK_DEV Fq12 conj() const { Fq12 r; r.c0 = c0; r.c1 = c1.neg(); return r; }   // the q^6 Frobenius (w -> -w)
    // slot t of sum_t c_t w^t:  t = 2 j + i  <->  (w^i, v^j)
    ZK_DEV Fq2 &slot(int t) { Fq6 &h = (t & 1) ? c1 : c0; int j = t >> 1; return j == 0 ? h.c0 : j == 1 ? h.c1 : h.c2; }
};
static ZK_PTFN Fq12 mul12(const Fq12 &a, const Fq12 &b) {
    Fq6 t0 = mul6(a.c0, b.c0), t1 = mul6(a.c1, b.c1);
    Fq12 r;
    r.c1 = mul6(a.c0 + a.c1, b.c0 + b.c1) - t0 - t1;
    r.c0 = t0 + t1.mul_v();
    return r;
}
static ZK_PTFN Fq12 sqr12(const Fq12 &a) {
    Fq6 ab = mul6(a.c0, a.c1);
    Fq12 r;
    r.c0 = mul6(a.c0 + a.c1, a.c0 + a.c1.mul_v()) - ab - ab.mul_v();
    r.c1 = ab + ab;
    return r;
}
static ZK_PTFN Fq12 inv12(const Fq12 &a) {   // 1 / (c0 + c1 w) = (c0 - c1 w) / (c0^2 - v c1^2)
    Fq6 n = inv6(mul6(a.c0, a.c0) - mul6(a.c1, a.c1).mul_v());
    Fq12 r; r.c0 = mul6(a.c0, n); r.c1 = mul6(a.c1, n).neg(); return r;
}
// f * (c2 + (c1) v + (c0) v w) with the Fq scalings of ell() already applied by the caller: 13 Fq2 products
static ZK_PTFN Fq12 mul12_014(const Fq12 &f, const Fq2 &s0, const Fq2 &s1, const Fq2 &s4) {
    Fq6 t0 = mul6_01(f.c0, s0, s1), t1 = mul6_1(f.c1, s4);
    Fq12 r;
    r.c1 = mul6_01(f.c0 + f.c1, s0, s1 + s4) - t0 - t1;
    r.c0 = t0 + t1.mul_v();
    return r;
}
static ZK_PTFN Fq12 frobenius12(const Fq12 &a, int k) {
    Fq2 g = frob_gamma(k), gp = g;
    Fq12 r = a;
    for (int t = 0; t < 6; t++) {
        Fq2 c = r.slot(t);
        if (k & 1) c = conj2(c);
        if (t == 1) c = c * g;
        else if (t > 1) { gp = gp * g; c = c * gp; }
        r.slot(t) = c;
    }
    return r;
}
// Squaring in the cyclotomic subgroup (Granger-Scott): view Fq12 as Fq4[w]/(w^3 - s), Fq4 = Fq2[s]/(s^2 - xi), s = w^3, so
// g = A + B w + C w^2 with A = (c_0, c_3), B = (c_1, c_4), C = (c_2, c_5) in slot numbering; for g of norm one over Fq6,
//   g^2 = (3 A^2 - 2 conj(A)) + (3 s C^2 + 2 conj(B)) w + (3 B^2 - 2 conj(C)) w^2        (9 Fq2 squarings instead of 12 products)
ZK_DEV void sqr4(const Fq2 &a, const Fq2 &b, Fq2 &r0, Fq2 &r1) {   // (a + b s)^2
    Fq2 aa = a.sqr(), bb = b.sqr();
    r1 = (a + b).sqr() - aa - bb;
    r0 = aa + mul_xi(bb);
}
static ZK_PTFN Fq12 cyclotomic_sqr(const Fq12 &g) {
    Fq2 t0, t1, u0, u1, v0, v1;
    sqr4(g.c0.c0, g.c1.c1, t0, t1);      // A^2
    sqr4(g.c1.c0, g.c0.c2, u0, u1);      // B^2
    sqr4(g.c0.c1, g.c1.c2, v0, v1);      // C^2
    Fq12 r;
    r.c0.c0 = (t0 - g.c0.c0).dbl() + t0;             // 3 t0 - 2 c_0
    r.c1.c1 = (t1 + g.c1.c1).dbl() + t1;             // 3 t1 + 2 c_3
    Fq2 sv = mul_xi(v1);                             // s C^2 = xi v1 + v0 s
    r.c1.c0 = (sv + g.c1.c0).dbl() + sv;             // 3 xi v1 + 2 c_1
    r.c0.c2 = (v0 - g.c0.c2).dbl() + v0;             // 3 v0 - 2 c_4
    r.c0.c1 = (u0 - g.c0.c1).dbl() + u0;             // 3 u0 - 2 c_2
    r.c1.c2 = (u1 + g.c1.c2).dbl() + u1;             // 3 u1 + 2 c_5
    return r;
}
// f^|x| by square-and-multiply, then conjugated: f^x for f in the cyclotomic subgroup
static ZK_PTFN Fq12 exp_x(const Fq12 &f) {
    Fq12 r = f;
    for (int i = 62; i >= 0; i--) {
        r = cyclotomic_sqr(r);
        if ((BLS_X_ABS >> i) & 1) r = mul12(r, f);
    }
    return r.conj();
}
// returns false when f == 0 (Engine::final_exponentiation -> None)
static ZK_PTFN bool final_exponentiation(const Fq12 &f, Fq12 &out) {
    if (f.is_zero()) return false;
    Fq12 g = mul12(f.conj(), inv12(f));                  // f^(q^6 - 1)
    g = mul12(frobenius12(g, 2), g);                      // ^(q^2 + 1): now in the cyclotomic subgroup, inverse = conj
    Fq12 a = mul12(exp_x(g), g.conj());                   // g^(x-1)
    a = mul12(exp_x(a), a.conj());                        // g^((x-1)^2)            = g^l3
    Fq12 b = exp_x(a);                                    // g^l2
    Fq12 c = mul12(exp_x(b), a.conj());                   // g^l1
    Fq12 d = mul12(exp_x(c), mul12(cyclotomic_sqr(g), g));  // g^l0
    out = mul12(mul12(d, frobenius12(c, 1)), mul12(frobenius12(b, 2), frobenius12(a, 3)));
    return true;
}

// ---- prepared G2 -------------------------------------------------------------------------------------------
struct LineCoeff { Fq2 c0, c1, c2; };
struct G2Jac { Fq2 x, y, z; };

static ZK_PTFN LineCoeff doubling_step(G2Jac &t) {
    Fq2 xx = t.x.sqr(), yy = t.y.sqr(), zz = t.z.sqr();
    Fq2 s = (t.x * yy).dbl().dbl();                       // 4 X Y^2
    Fq2 e = xx.dbl() + xx;                                // 3 X^2
    Fq2 x3 = e.sqr() - s.dbl();
    Fq2 z3 = (t.y * t.z).dbl();
    Fq2 y3 = e * (s - x3) - yy.sqr().dbl().dbl().dbl();
    LineCoeff l;
    l.c0 = (z3 * zz).dbl();
    l.c1 = (e * zz).dbl().neg();
    l.c2 = (e * t.x).dbl() - yy.dbl().dbl();
    t.x = x3; t.y = y3; t.z = z3;
    return l;
}
static ZK_PTFN LineCoeff addition_step(G2Jac &t, const Affine<Fq2> &q) {
    Fq2 zz = t.z.sqr();
    Fq2 h = q.x * zz - t.x;
    Fq2 r = (q.y * (t.z * zz) - t.y).dbl();
    Fq2 hh = h.sqr();
    Fq2 z3 = (t.z * h).dbl();
    Fq2 v4 = (t.x * hh).dbl().dbl();
    Fq2 h34 = (h * hh).dbl().dbl();
    Fq2 x3 = r.sqr() - h34 - v4.dbl();
    Fq2 y3 = r * (v4 - x3) - (t.y * h34).dbl();
    LineCoeff l;
    l.c0 = z3.dbl();
    l.c1 = r.dbl().neg();
    l.c2 = (r * q.x - q.y * z3).dbl();
    t.x = x3; t.y = y3; t.z = z3;
    return l;
}
// coefficient order = consumption order of the Miller loop; `stride` separates consecutive coefficients in `out`
static ZK_PTFN void g2_prepare(const Affine<Fq2> &q, LineCoeff *out, size_t stride) {
    G2Jac t; t.x = q.x; t.y = q.y; t.z = Fq2::one();
    int n = 0;
    for (int i = 62; i >= 1; i--) {
        out[(size_t)(n++) * stride] = doubling_step(t);
        if ((BLS_X_ABS >> i) & 1) out[(size_t)(n++) * stride] = addition_step(t, q);
    }
    out[(size_t)(n++) * stride] = doubling_step(t);
}
ZK_DEV Fq12 ell(const Fq12 &f, const LineCoeff &c, const Affine<Fq> &p) {
    return mul12_014(f, c.c2, mul_fq(c.c1, p.x), mul_fq(c.c0, p.y));
}
// one pair; infinity on either side contributes 1 (mod.rs:50-54)
static ZK_PTFN Fq12 miller_loop(const Affine<Fq> &p, const LineCoeff *coeffs, size_t stride, bool g2_inf) {
    Fq12 f = Fq12::one();
    if (p.is_inf() || g2_inf) return f;
    int n = 0;
    for (int i = 62; i >= 1; i--) {
        f = ell(f, coeffs[(size_t)(n++) * stride], p);
        if ((BLS_X_ABS >> i) & 1) f = ell(f, coeffs[(size_t)(n++) * stride], p);
        f = sqr12(f);
    }
    f = ell(f, coeffs[(size_t)(n++) * stride], p);
    return f.conj();
}

}  // namespace zkpair
