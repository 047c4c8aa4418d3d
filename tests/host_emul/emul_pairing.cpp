// CPU unit-test harness for zero_chain_b200/csrc/pairing.cuh (same source as the device build, ZK_HOST_EMUL carry
// primitives).  Test infrastructure only — never linked into libzkb200.so.
#define ZK_HOST_EMUL 1
#include "codec.cuh"
#include "pairing.cuh"
#include <string.h>
#include <vector>
using namespace zkpair;
static Fq12 ld12(const uint32_t *a) { Fq12 x; memcpy(&x, a, sizeof(x)); return x; }
static void st12(uint32_t *o, const Fq12 &x) { memcpy(o, &x, sizeof(x)); }
extern "C" {
int emu_sizeof_fq12() { return (int)sizeof(Fq12); }
int emu_sizeof_coeff() { return (int)sizeof(LineCoeff); }
void emu_f12_mul(const uint32_t *a, const uint32_t *b, uint32_t *o) { st12(o, mul12(ld12(a), ld12(b))); }
void emu_f12_cyc_sqr(const uint32_t *a, uint32_t *o) { st12(o, cyclotomic_sqr(ld12(a))); }
void emu_f12_sqr(const uint32_t *a, uint32_t *o) { st12(o, sqr12(ld12(a))); }
void emu_f12_inv(const uint32_t *a, uint32_t *o) { st12(o, inv12(ld12(a))); }
void emu_f12_frob(const uint32_t *a, int k, uint32_t *o) { st12(o, frobenius12(ld12(a), k)); }
int emu_final_exp(const uint32_t *a, uint32_t *o) { Fq12 r = Fq12::one(); bool ok = final_exponentiation(ld12(a), r); st12(o, r); return ok; }
void emu_g2_prepare(const uint32_t *q, uint32_t *out) {
    Affine<Fq2> p; memcpy(&p, q, sizeof(p));
    std::vector<LineCoeff> c(N_COEFFS);
    g2_prepare(p, c.data(), 1);
    memcpy(out, c.data(), sizeof(LineCoeff) * N_COEFFS);
}
void emu_miller(const uint32_t *p, const uint32_t *coeffs, uint32_t *o) {
    Affine<Fq> g; memcpy(&g, p, sizeof(g));
    std::vector<LineCoeff> c(N_COEFFS);
    memcpy(c.data(), coeffs, sizeof(LineCoeff) * N_COEFFS);
    st12(o, miller_loop(g, c.data(), 1, false));
}
// Compressed::into_affine: returns the DEC_* code; out = affine Montgomery limbs
int emu_decode_g1c(const uint8_t *in, uint32_t *out) { Affine<Fq> p = Affine<Fq>::inf(); int e = zkcodec::decode_compressed(p, in); memcpy(out, &p, sizeof(p)); return e; }
int emu_decode_g2c(const uint8_t *in, uint32_t *out) { Affine<Fq2> p = Affine<Fq2>::inf(); int e = zkcodec::decode_compressed(p, in); memcpy(out, &p, sizeof(p)); return e; }
// subgroup tests on affine Montgomery limbs: bit 0 = endomorphism test, bit 1 = multiplication by r
int emu_g1_subgroup(const uint32_t *in) { Affine<Fq> p; memcpy(&p, in, sizeof(p)); return (int)zkcodec::in_subgroup(p) | ((int)zkcodec::in_subgroup_by_order(p) << 1); }
int emu_g2_subgroup(const uint32_t *in) { Affine<Fq2> p; memcpy(&p, in, sizeof(p)); return (int)zkcodec::in_subgroup(p) | ((int)zkcodec::in_subgroup_by_order(p) << 1); }
}
