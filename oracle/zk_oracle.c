/* =====================================================================================
 * TEST INFRASTRUCTURE — CPU oracle for the Groth16 prover hot path (plain C99 + OpenMP).
 *
 * This file is NOT part of the product.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call it.  The product
 * (libzkb200.so) never links or dlopens it and fails loudly without its CUDA kernels.
 *
 * What it restates (reference file:line under LayerXcom/zero-chain):
 *   Fq / Fr / Fq2 arithmetic      core/pairing/src/bls12_381/{fq.rs,fr.rs,fq2.rs}  (see field_tmpl.inc)
 *   G1 / G2 Jacobian group law    core/pairing/src/bls12_381/ec.rs:224-618         (see curve_tmpl.inc)
 *   point encodings               core/pairing/src/bls12_381/ec.rs:686-867, 1343-1549
 *   Proof encoding                core/bellman-verifier/src/lib.rs:55-65
 *   prover call site              core/proofs/src/confidential.rs:149 (create_random_proof)
 * and, because the prover itself is an un-vendored dependency (bellman 0.1.0 @
 * LayerXcom/librustzcash#2c19687150cd5daddb793e0c8e651a95f10d8a21, Cargo.lock:210-212), the
 * published bellman algorithms, restated from SURVEY.md §3.2/§3.3:
 *   multiexp (Pippenger, c = 3 if n < 32 else ceil(ln n), 0/1 fast paths, density maps)
 *   EvaluationDomain (radix-2 DIT fft/ifft/coset_fft/icoset_fft/divide_by_z_on_coset)
 *   groth16::create_proof, groth16::Parameters::{read,write}
 *
 * PARITY PIN: the arithmetic, group law and encodings are pinned by the reference's own
 * known-answer vectors (tests/golden/kats.json + the 4x1000-point encoding files, see
 * tests/test_oracle_kats.py).  PROVER OUTPUT BYTES ARE UNPINNED by the reference ("parity
 * unpinned": no reference test fixes r,s and asserts proof bytes) — the prover
 * restatement is instead checked against closed-form trapdoor algebra in oracle/pyref.py.
 * ===================================================================================== */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

typedef struct { uint64_t l[6]; } fq_t;
typedef struct { uint64_t l[4]; } fr_t;

/* ---- constants (fq.rs:5-43, fr.rs:4-55) ------------------------------------------- */
static const uint64_t FQ_MODULUS[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                       0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t FQ_R[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                                 0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
static const uint64_t FQ_R2[6] = {0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                                  0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL};
#define FQ_INV 0x89f3fffcfffcfffdULL
static const uint64_t FR_MODULUS[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                       0x73eda753299d7d48ULL};
static const uint64_t FR_R[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL,
                                 0x1824b159acc5056fULL};
static const uint64_t FR_R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL,
                                  0x0748d9d99f59ff11ULL};
#define FR_INV 0xfffffffeffffffffULL
#define FR_S 32
/* GENERATOR = 7, ROOT_OF_UNITY = 7^t (Montgomery form; fr.rs:38-55) */
static const uint64_t FR_GENERATOR[4] = {0x0000000efffffff1ULL, 0x17e363d300189c0fULL, 0xff9c57876f8457b0ULL,
                                         0x351332208fc5a8c4ULL};
static const uint64_t FR_ROOT_OF_UNITY[4] = {0xb9b58d8c5f0e466aULL, 0x5b1b4c801819d7ecULL, 0x0af53ae352a31e64ULL,
                                             0x5bf3adda19e9b27bULL};

/* ---- Fq ----------------------------------------------------------------------------- */
#define FN(x) CAT(fq_, x)
#define FT fq_t
#define NL 6
#define F_MODULUS FQ_MODULUS
#define F_R FQ_R
#define F_R2 FQ_R2
#define F_INV FQ_INV
#include "field_tmpl.inc"
#undef FN
#undef FT
#undef NL
#undef F_MODULUS
#undef F_R
#undef F_R2
#undef F_INV
/* ---- Fr ----------------------------------------------------------------------------- */
#define FN(x) CAT(fr_, x)
#define FT fr_t
#define NL 4
#define F_MODULUS FR_MODULUS
#define F_R FR_R
#define F_R2 FR_R2
#define F_INV FR_INV
#include "field_tmpl.inc"
#undef FN
#undef FT
#undef NL
#undef F_MODULUS
#undef F_R
#undef F_R2
#undef F_INV

/* ---- Fq2 = Fq[u]/(u^2+1)  (fq2.rs: square 109-123, mul 145-158, inverse 183-201) ------ */
typedef struct { fq_t c0, c1; } fq2_t;
static inline int fq2_is_zero(const fq2_t *a) { return fq_is_zero(&a->c0) && fq_is_zero(&a->c1); }
static inline int fq2_eq(const fq2_t *a, const fq2_t *b) { return fq_eq(&a->c0, &b->c0) && fq_eq(&a->c1, &b->c1); }
static inline void fq2_set_zero(fq2_t *a) { fq_set_zero(&a->c0); fq_set_zero(&a->c1); }
static inline void fq2_set_one(fq2_t *a) { fq_set_one(&a->c0); fq_set_zero(&a->c1); }
static inline void fq2_add(fq2_t *r, const fq2_t *a, const fq2_t *b) { fq_add(&r->c0, &a->c0, &b->c0); fq_add(&r->c1, &a->c1, &b->c1); }
static inline void fq2_sub(fq2_t *r, const fq2_t *a, const fq2_t *b) { fq_sub(&r->c0, &a->c0, &b->c0); fq_sub(&r->c1, &a->c1, &b->c1); }
static inline void fq2_dbl(fq2_t *r, const fq2_t *a) { fq_dbl(&r->c0, &a->c0); fq_dbl(&r->c1, &a->c1); }
static inline void fq2_neg(fq2_t *r, const fq2_t *a) { fq_neg(&r->c0, &a->c0); fq_neg(&r->c1, &a->c1); }
static inline void fq2_mul(fq2_t *r, const fq2_t *a, const fq2_t *b) {
    /* Karatsuba: aa = a0 b0, bb = a1 b1, o = (b0+b1); c1 = (a0+a1) o - aa - bb; c0 = aa - bb */
    fq_t aa, bb, o, t;
    fq_mul(&aa, &a->c0, &b->c0);
    fq_mul(&bb, &a->c1, &b->c1);
    fq_add(&o, &b->c0, &b->c1);
    fq_add(&t, &a->c1, &a->c0);
    fq_mul(&t, &t, &o);
    fq_sub(&t, &t, &aa);
    fq_sub(&r->c1, &t, &bb);
    fq_sub(&r->c0, &aa, &bb);
}
static inline void fq2_sqr(fq2_t *r, const fq2_t *a) {
    /* complex squaring: ab = a0 a1; c0 = (a0+a1)(a0-a1) ... fq2.rs:109-123 */
    fq_t ab, c0c1, c0;
    fq_mul(&ab, &a->c0, &a->c1);
    fq_add(&c0c1, &a->c0, &a->c1);
    fq_neg(&c0, &a->c1);
    fq_add(&c0, &c0, &a->c0);
    fq_mul(&c0, &c0, &c0c1);
    fq_sub(&c0, &c0, &ab);
    fq_add(&c0, &c0, &ab);
    r->c0 = c0;
    fq_dbl(&r->c1, &ab);
}
static inline int fq2_inv(fq2_t *r, const fq2_t *a) {
    fq_t t0, t1;
    fq_sqr(&t1, &a->c1);
    fq_sqr(&t0, &a->c0);
    fq_add(&t0, &t0, &t1);
    if (fq_inv(&t0, &t0)) return -1;
    fq_mul(&r->c0, &a->c0, &t0);
    fq_mul(&t1, &a->c1, &t0);
    fq_neg(&r->c1, &t1);
    return 0;
}

/* ---- curves ---------------------------------------------------------------------------- */
typedef struct { fq_t x, y; int inf; } g1_aff;
typedef struct { fq_t x, y, z; } g1_jac;
typedef struct { fq2_t x, y; int inf; } g2_aff;
typedef struct { fq2_t x, y, z; } g2_jac;

/* B_COEFF = 4 (Montgomery; fq.rs:69-76); B for G2 = 4(1+u) (ec.rs:1567-1572) */
static const fq_t G1_B = {{0xaa270000000cfff3ULL, 0x53cc0032fc34000aULL, 0x478fe97a6b0a807fULL,
                           0xb1d37ebee6ba24d7ULL, 0x8ec9733bbf78ab2fULL, 0x09d645513d83de7eULL}};
static const fq2_t G2_B = {{{0xaa270000000cfff3ULL, 0x53cc0032fc34000aULL, 0x478fe97a6b0a807fULL,
                             0xb1d37ebee6ba24d7ULL, 0x8ec9733bbf78ab2fULL, 0x09d645513d83de7eULL}},
                           {{0xaa270000000cfff3ULL, 0x53cc0032fc34000aULL, 0x478fe97a6b0a807fULL,
                             0xb1d37ebee6ba24d7ULL, 0x8ec9733bbf78ab2fULL, 0x09d645513d83de7eULL}}};

#define CN(x) CAT(g1_, x)
#define BF(x) CAT(fq_, x)
#define BT fq_t
#define AFF g1_aff
#define JAC g1_jac
#define C_B (&G1_B)
#include "curve_tmpl.inc"
#undef CN
#undef BF
#undef BT
#undef AFF
#undef JAC
#undef C_B
#define CN(x) CAT(g2_, x)
#define BF(x) CAT(fq2_, x)
#define BT fq2_t
#define AFF g2_aff
#define JAC g2_jac
#define C_B (&G2_B)
#include "curve_tmpl.inc"
#undef CN
#undef BF
#undef BT
#undef AFF
#undef JAC
#undef C_B

/* generators (Montgomery; fq.rs:85-136) */
static const g1_aff G1_GEN = {
    {{0x5cb38790fd530c16ULL, 0x7817fc679976fff5ULL, 0x154f95c7143ba1c1ULL, 0xf0ae6acdf3d0e747ULL, 0xedce6ecc21dbf440ULL, 0x120177419e0bfb75ULL}},
    {{0xbaac93d50ce72271ULL, 0x8c22631a7918fd8eULL, 0xdd595f13570725ceULL, 0x51ac582950405194ULL, 0x0e1c8c3fad0059c0ULL, 0x0bbc3efc5008a26aULL}}, 0};
static const g2_aff G2_GEN = {
    {{{0xf5f28fa202940a10ULL, 0xb3f5fb2687b4961aULL, 0xa1a893b53e2ae580ULL, 0x9894999d1a3caee9ULL, 0x6f67b7631863366bULL, 0x058191924350bcd7ULL}},
     {{0xa5a9c0759e23f606ULL, 0xaaa0c59dbccd60c3ULL, 0x3bb17e18e2867806ULL, 0x1b1ab6cc8541b367ULL, 0xc2b6ed0ef2158547ULL, 0x11922a097360edf3ULL}}},
    {{{0x4c730af860494c4aULL, 0x597cfa1f5e369c5aULL, 0xe7e6856caa0a635aULL, 0xbbefb5e96e0d495fULL, 0x07d3a975f0ef25a2ULL, 0x0083fd8e7e80dae5ULL}},
     {{0xadc0fc92df64b05dULL, 0x18aa270a2b1461dcULL, 0x86adac6a3be4eba0ULL, 0x79495c4ec93da33aULL, 0xe7175850a43ccaedULL, 0x0b2bc2a163de1bf2ULL}}}, 0};

/* =========================================================================================
 * flat layouts used across the ctypes boundary
 *   Fq  : 6 LE u64 limbs, Montgomery      Fr : 4 LE u64 limbs (Montgomery or canonical as stated)
 *   G1 affine "limb form": x[6] y[6]  (96 B), infinity = all-zero
 *   G2 affine "limb form": x.c0[6] x.c1[6] y.c0[6] y.c1[6] (192 B), infinity = all-zero
 * ========================================================================================= */
static void g1_load(g1_aff *p, const uint64_t *w) {
    memcpy(p->x.l, w, 48); memcpy(p->y.l, w + 6, 48);
    p->inf = fq_is_zero(&p->x) && fq_is_zero(&p->y);
    if (p->inf) g1_aff_set_zero(p);
}
static void g1_store(uint64_t *w, const g1_aff *p) {
    if (p->inf) { memset(w, 0, 96); return; }
    memcpy(w, p->x.l, 48); memcpy(w + 6, p->y.l, 48);
}
static void g2_load(g2_aff *p, const uint64_t *w) {
    memcpy(&p->x, w, 96); memcpy(&p->y, w + 12, 96);
    p->inf = fq2_is_zero(&p->x) && fq2_is_zero(&p->y);
    if (p->inf) g2_aff_set_zero(p);
}
static void g2_store(uint64_t *w, const g2_aff *p) {
    if (p->inf) { memset(w, 0, 192); return; }
    memcpy(w, &p->x, 96); memcpy(w + 12, &p->y, 96);
}

/* ---- field-level exports (KAT harness) -------------------------------------------------- */
EXPORT void zko_fq_mul(const uint64_t *a, const uint64_t *b, uint64_t *o) { fq_mul((fq_t *)o, (const fq_t *)a, (const fq_t *)b); }
EXPORT void zko_fq_sqr(const uint64_t *a, uint64_t *o) { fq_sqr((fq_t *)o, (const fq_t *)a); }
EXPORT void zko_fq_add(const uint64_t *a, const uint64_t *b, uint64_t *o) { fq_add((fq_t *)o, (const fq_t *)a, (const fq_t *)b); }
EXPORT void zko_fq_sub(const uint64_t *a, const uint64_t *b, uint64_t *o) { fq_sub((fq_t *)o, (const fq_t *)a, (const fq_t *)b); }
EXPORT void zko_fq_neg(const uint64_t *a, uint64_t *o) { fq_neg((fq_t *)o, (const fq_t *)a); }
EXPORT int zko_fq_inv(const uint64_t *a, uint64_t *o) { return fq_inv((fq_t *)o, (const fq_t *)a); }
EXPORT int zko_fq_from_repr(const uint64_t *a, uint64_t *o) { return fq_from_repr((fq_t *)o, a); }
EXPORT void zko_fq_into_repr(const uint64_t *a, uint64_t *o) { fq_into_repr(o, (const fq_t *)a); }
EXPORT void zko_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t *o) { fr_mul((fr_t *)o, (const fr_t *)a, (const fr_t *)b); }
EXPORT void zko_fr_sqr(const uint64_t *a, uint64_t *o) { fr_sqr((fr_t *)o, (const fr_t *)a); }
EXPORT void zko_fr_add(const uint64_t *a, const uint64_t *b, uint64_t *o) { fr_add((fr_t *)o, (const fr_t *)a, (const fr_t *)b); }
EXPORT void zko_fr_sub(const uint64_t *a, const uint64_t *b, uint64_t *o) { fr_sub((fr_t *)o, (const fr_t *)a, (const fr_t *)b); }
EXPORT void zko_fr_neg(const uint64_t *a, uint64_t *o) { fr_neg((fr_t *)o, (const fr_t *)a); }
EXPORT int zko_fr_inv(const uint64_t *a, uint64_t *o) { return fr_inv((fr_t *)o, (const fr_t *)a); }
EXPORT int zko_fr_from_repr(const uint64_t *a, uint64_t *o) { return fr_from_repr((fr_t *)o, a); }
EXPORT void zko_fr_into_repr(const uint64_t *a, uint64_t *o) { fr_into_repr(o, (const fr_t *)a); }
EXPORT void zko_fq2_mul(const uint64_t *a, const uint64_t *b, uint64_t *o) { fq2_t r; fq2_mul(&r, (const fq2_t *)a, (const fq2_t *)b); memcpy(o, &r, 96); }
EXPORT void zko_fq2_sqr(const uint64_t *a, uint64_t *o) { fq2_t r; fq2_sqr(&r, (const fq2_t *)a); memcpy(o, &r, 96); }
EXPORT int zko_fq2_inv(const uint64_t *a, uint64_t *o) { fq2_t r; int e = fq2_inv(&r, (const fq2_t *)a); memcpy(o, &r, 96); return e; }
/* n independent Montgomery products (bench helper + witness generation: c = a∘b) */
EXPORT void zko_fr_mul_many(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) fr_mul((fr_t *)(o + 4 * i), (const fr_t *)(a + 4 * i), (const fr_t *)(b + 4 * i));
}
EXPORT void zko_fr_from_repr_many(const uint64_t *a, uint64_t *o, size_t n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) fr_from_repr((fr_t *)(o + 4 * i), a + 4 * i);
}
EXPORT void zko_fr_into_repr_many(const uint64_t *a, uint64_t *o, size_t n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) fr_into_repr(o + 4 * i, (const fr_t *)(a + 4 * i));
}
EXPORT double zko_fq_mul_bench(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t iters) {
    fq_t x = *(const fq_t *)a, y = *(const fq_t *)b;
    for (size_t i = 0; i < iters; i++) { fq_mul(&x, &x, &y); }
    memcpy(o, &x, 48);
    return (double)iters;
}

/* ---- point-level exports ---------------------------------------------------------------- */
EXPORT void zko_g1_generator(uint64_t *o) { g1_store(o, &G1_GEN); }
EXPORT void zko_g2_generator(uint64_t *o) { g2_store(o, &G2_GEN); }
EXPORT void zko_g1_add(const uint64_t *a, const uint64_t *b, uint64_t *o) {
    g1_aff pa, pb, r; g1_jac j, k;
    g1_load(&pa, a); g1_load(&pb, b); g1_from_affine(&j, &pa); g1_from_affine(&k, &pb);
    g1_add(&j, &k); g1_into_affine(&r, &j); g1_store(o, &r);
}
EXPORT void zko_g1_add_mixed(const uint64_t *a, const uint64_t *b, uint64_t *o) {
    g1_aff pa, pb, r; g1_jac j;
    g1_load(&pa, a); g1_load(&pb, b); g1_from_affine(&j, &pa);
    g1_add_mixed(&j, &pb); g1_into_affine(&r, &j); g1_store(o, &r);
}
EXPORT void zko_g1_double(const uint64_t *a, uint64_t *o) {
    g1_aff pa, r; g1_jac j; g1_load(&pa, a); g1_from_affine(&j, &pa); g1_dbl(&j); g1_into_affine(&r, &j); g1_store(o, &r);
}
EXPORT void zko_g1_mul(const uint64_t *a, const uint64_t *k, uint64_t *o) {
    g1_aff pa, r; g1_jac j; g1_load(&pa, a); g1_aff_mul(&j, &pa, k); g1_into_affine(&r, &j); g1_store(o, &r);
}
EXPORT int zko_g1_check(const uint64_t *a) {   /* bit0 on-curve, bit1 in-subgroup */
    g1_aff pa; g1_load(&pa, a); int oc = g1_is_on_curve(&pa); return oc | ((oc && g1_in_subgroup(&pa)) << 1);
}
EXPORT void zko_g2_add(const uint64_t *a, const uint64_t *b, uint64_t *o) {
    g2_aff pa, pb, r; g2_jac j, k;
    g2_load(&pa, a); g2_load(&pb, b); g2_from_affine(&j, &pa); g2_from_affine(&k, &pb);
    g2_add(&j, &k); g2_into_affine(&r, &j); g2_store(o, &r);
}
EXPORT void zko_g2_add_mixed(const uint64_t *a, const uint64_t *b, uint64_t *o) {
    g2_aff pa, pb, r; g2_jac j;
    g2_load(&pa, a); g2_load(&pb, b); g2_from_affine(&j, &pa);
    g2_add_mixed(&j, &pb); g2_into_affine(&r, &j); g2_store(o, &r);
}
EXPORT void zko_g2_double(const uint64_t *a, uint64_t *o) {
    g2_aff pa, r; g2_jac j; g2_load(&pa, a); g2_from_affine(&j, &pa); g2_dbl(&j); g2_into_affine(&r, &j); g2_store(o, &r);
}
EXPORT void zko_g2_mul(const uint64_t *a, const uint64_t *k, uint64_t *o) {
    g2_aff pa, r; g2_jac j; g2_load(&pa, a); g2_aff_mul(&j, &pa, k); g2_into_affine(&r, &j); g2_store(o, &r);
}
EXPORT int zko_g2_check(const uint64_t *a) {
    g2_aff pa; g2_load(&pa, a); int oc = g2_is_on_curve(&pa); return oc | ((oc && g2_in_subgroup(&pa)) << 1);
}

/* ---- encodings (ec.rs:686-867, 1343-1549; PrimeFieldRepr::write_be lib.rs:408-431) -------- */
static void fq_write_be(uint8_t *out, const fq_t *a) {
    uint64_t r[6]; fq_into_repr(r, a);
    for (int i = 0; i < 6; i++) for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(r[5 - i] >> (56 - 8 * b));
}
static int fq_read_be(fq_t *a, const uint8_t *in, uint8_t mask0) {
    uint64_t r[6];
    for (int i = 0; i < 6; i++) {
        uint64_t v = 0;
        for (int b = 0; b < 8; b++) { uint8_t c = in[8 * i + b]; if (i == 0 && b == 0) c &= mask0; v = (v << 8) | c; }
        r[5 - i] = v;
    }
    return fq_from_repr(a, r);
}
static int fq_lex_gt_neg(const fq_t *y) {   /* y > -y on canonical integers (fq.rs:708-713) */
    fq_t n; fq_neg(&n, y);
    uint64_t a[6], b[6]; fq_into_repr(a, y); fq_into_repr(b, &n);
    for (int i = 5; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    return 0;
}
static int fq2_lex_gt_neg(const fq2_t *y) { /* Fq2 order: c1 first, then c0 (fq2.rs:21-30) */
    fq2_t n; fq2_neg(&n, y);
    uint64_t a[6], b[6];
    fq_into_repr(a, &y->c1); fq_into_repr(b, &n.c1);
    for (int i = 5; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    fq_into_repr(a, &y->c0); fq_into_repr(b, &n.c0);
    for (int i = 5; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    return 0;
}
static void g1_encode(uint8_t *out, const g1_aff *p, int compressed) {
    int len = compressed ? 48 : 96;
    memset(out, 0, len);
    if (p->inf) out[0] |= 0x40;
    else {
        fq_write_be(out, &p->x);
        if (!compressed) fq_write_be(out + 48, &p->y);
        else if (fq_lex_gt_neg(&p->y)) out[0] |= 0x20;
    }
    if (compressed) out[0] |= 0x80;
}
static void g2_encode(uint8_t *out, const g2_aff *p, int compressed) {
    int len = compressed ? 96 : 192;
    memset(out, 0, len);
    if (p->inf) out[0] |= 0x40;
    else {
        fq_write_be(out, &p->x.c1); fq_write_be(out + 48, &p->x.c0);
        if (!compressed) { fq_write_be(out + 96, &p->y.c1); fq_write_be(out + 144, &p->y.c0); }
        else if (fq2_lex_gt_neg(&p->y)) out[0] |= 0x20;
    }
    if (compressed) out[0] |= 0x80;
}
/* error codes: 0 ok, 1 UnexpectedCompressionMode, 2 UnexpectedInformation, 3 CoordinateDecodingError,
 *              4 NotOnCurve, 5 NotInSubgroup  (GroupDecodingError, core/pairing/src/lib.rs) */
static int g1_decode_uncompressed(g1_aff *p, const uint8_t *in, int checked) {
    if (in[0] & 0x80) return 1;
    if (in[0] & 0x40) {
        if (in[0] & 0x3f) return 2;
        for (int i = 1; i < 96; i++) if (in[i]) return 2;
        g1_aff_set_zero(p); return 0;
    }
    if (in[0] & 0x20) return 2;
    if (fq_read_be(&p->x, in, 0x1f) || fq_read_be(&p->y, in + 48, 0xff)) return 3;
    p->inf = 0;
    if (checked) { if (!g1_is_on_curve(p)) return 4; if (!g1_in_subgroup(p)) return 5; }
    return 0;
}
static int g2_decode_uncompressed(g2_aff *p, const uint8_t *in, int checked) {
    if (in[0] & 0x80) return 1;
    if (in[0] & 0x40) {
        if (in[0] & 0x3f) return 2;
        for (int i = 1; i < 192; i++) if (in[i]) return 2;
        g2_aff_set_zero(p); return 0;
    }
    if (in[0] & 0x20) return 2;
    if (fq_read_be(&p->x.c1, in, 0x1f) || fq_read_be(&p->x.c0, in + 48, 0xff) ||
        fq_read_be(&p->y.c1, in + 96, 0xff) || fq_read_be(&p->y.c0, in + 144, 0xff)) return 3;
    p->inf = 0;
    if (checked) { if (!g2_is_on_curve(p)) return 4; if (!g2_in_subgroup(p)) return 5; }
    return 0;
}
EXPORT void zko_g1_encode(const uint64_t *a, int compressed, uint8_t *out) { g1_aff p; g1_load(&p, a); g1_encode(out, &p, compressed); }
EXPORT void zko_g2_encode(const uint64_t *a, int compressed, uint8_t *out) { g2_aff p; g2_load(&p, a); g2_encode(out, &p, compressed); }
EXPORT int zko_g1_decode_uncompressed(const uint8_t *in, int checked, uint64_t *o) { g1_aff p; int e = g1_decode_uncompressed(&p, in, checked); if (!e) g1_store(o, &p); return e; }
EXPORT int zko_g2_decode_uncompressed(const uint8_t *in, int checked, uint64_t *o) { g2_aff p; int e = g2_decode_uncompressed(&p, in, checked); if (!e) g2_store(o, &p); return e; }
/* bulk decode of n uncompressed points into limb form (parallel) */
EXPORT int zko_g1_decode_many(const uint8_t *in, size_t n, int checked, uint64_t *o) {
    int err = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long)n; i++) { g1_aff p; int e = g1_decode_uncompressed(&p, in + 96 * i, checked); if (e) err = e; else g1_store(o + 12 * i, &p); }
    return err;
}
EXPORT int zko_g2_decode_many(const uint8_t *in, size_t n, int checked, uint64_t *o) {
    int err = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long)n; i++) { g2_aff p; int e = g2_decode_uncompressed(&p, in + 192 * i, checked); if (e) err = e; else g2_store(o + 24 * i, &p); }
    return err;
}

/* ---- fixed-base generation helpers (test-vector / toy-CRS construction) ----------------------
 * out[i] = scalars[i] * G  for canonical scalars; 8-bit fixed windows over a precomputed table.
 * out_enc: n * 96 (G1) or n * 192 (G2) bytes, uncompressed encoding; limb form if enc == 0. */
static void g1_batch_to_affine(g1_aff *out, g1_jac *in, long n) {
    /* Montgomery batch inversion (ec.rs:246-294 batch_normalization computes the same values) */
    fq_t *pre = (fq_t *)malloc(sizeof(fq_t) * (n + 1));
    fq_t acc; fq_set_one(&acc);
    for (long i = 0; i < n; i++) { pre[i] = acc; if (!g1_is_zero(&in[i])) fq_mul(&acc, &acc, &in[i].z); }
    fq_t inv; fq_inv(&inv, &acc);
    for (long i = n - 1; i >= 0; i--) {
        if (g1_is_zero(&in[i])) { g1_aff_set_zero(&out[i]); continue; }
        fq_t zi, zi2; fq_mul(&zi, &inv, &pre[i]); fq_mul(&inv, &inv, &in[i].z);
        fq_sqr(&zi2, &zi); fq_mul(&out[i].x, &in[i].x, &zi2); fq_mul(&zi2, &zi2, &zi); fq_mul(&out[i].y, &in[i].y, &zi2);
        out[i].inf = 0;
    }
    free(pre);
}
static void g2_batch_to_affine(g2_aff *out, g2_jac *in, long n) {
    fq2_t *pre = (fq2_t *)malloc(sizeof(fq2_t) * (n + 1));
    fq2_t acc; fq2_set_one(&acc);
    for (long i = 0; i < n; i++) { pre[i] = acc; if (!g2_is_zero(&in[i])) fq2_mul(&acc, &acc, &in[i].z); }
    fq2_t inv; fq2_inv(&inv, &acc);
    for (long i = n - 1; i >= 0; i--) {
        if (g2_is_zero(&in[i])) { g2_aff_set_zero(&out[i]); continue; }
        fq2_t zi, zi2; fq2_mul(&zi, &inv, &pre[i]); fq2_mul(&inv, &inv, &in[i].z);
        fq2_sqr(&zi2, &zi); fq2_mul(&out[i].x, &in[i].x, &zi2); fq2_mul(&zi2, &zi2, &zi); fq2_mul(&out[i].y, &in[i].y, &zi2);
        out[i].inf = 0;
    }
    free(pre);
}
EXPORT void zko_g1_fixed_base_many(const uint64_t *base, const uint64_t *scalars, size_t n, int enc, uint8_t *out) {
    g1_aff b; g1_load(&b, base);
    /* table[w][d] = d * 256^w * B, d = 1..255, affine */
    g1_jac *tj = (g1_jac *)malloc(sizeof(g1_jac) * 32 * 255);
    g1_aff *ta = (g1_aff *)malloc(sizeof(g1_aff) * 32 * 255);
    g1_jac cur; g1_from_affine(&cur, &b);
    for (int w = 0; w < 32; w++) {
        g1_jac run = cur;
        for (int d = 1; d <= 255; d++) { tj[w * 255 + d - 1] = run; g1_add(&run, &cur); }
        cur = run;   /* 256 * cur */
    }
    g1_batch_to_affine(ta, tj, 32 * 255);
    free(tj);
#pragma omp parallel
    {
        const long CH = 256;
        g1_jac *acc = (g1_jac *)malloc(sizeof(g1_jac) * CH);
        g1_aff *aff = (g1_aff *)malloc(sizeof(g1_aff) * CH);
#pragma omp for schedule(dynamic)
        for (long c0 = 0; c0 < (long)n; c0 += CH) {
            long m = (long)n - c0 < CH ? (long)n - c0 : CH;
            for (long i = 0; i < m; i++) {
                g1_set_zero(&acc[i]);
                const uint8_t *k = (const uint8_t *)(scalars + 4 * (c0 + i));
                for (int w = 0; w < 32; w++) if (k[w]) g1_add_mixed(&acc[i], &ta[w * 255 + k[w] - 1]);
            }
            g1_batch_to_affine(aff, acc, m);
            for (long i = 0; i < m; i++) {
                if (enc) g1_encode(out + 96 * (c0 + i), &aff[i], 0); else g1_store((uint64_t *)(out + 96 * (c0 + i)), &aff[i]);
            }
        }
        free(acc); free(aff);
    }
    free(ta);
}
EXPORT void zko_g2_fixed_base_many(const uint64_t *base, const uint64_t *scalars, size_t n, int enc, uint8_t *out) {
    g2_aff b; g2_load(&b, base);
    g2_jac *tj = (g2_jac *)malloc(sizeof(g2_jac) * 32 * 255);
    g2_aff *ta = (g2_aff *)malloc(sizeof(g2_aff) * 32 * 255);
    g2_jac cur; g2_from_affine(&cur, &b);
    for (int w = 0; w < 32; w++) {
        g2_jac run = cur;
        for (int d = 1; d <= 255; d++) { tj[w * 255 + d - 1] = run; g2_add(&run, &cur); }
        cur = run;
    }
    g2_batch_to_affine(ta, tj, 32 * 255);
    free(tj);
#pragma omp parallel
    {
        const long CH = 256;
        g2_jac *acc = (g2_jac *)malloc(sizeof(g2_jac) * CH);
        g2_aff *aff = (g2_aff *)malloc(sizeof(g2_aff) * CH);
#pragma omp for schedule(dynamic)
        for (long c0 = 0; c0 < (long)n; c0 += CH) {
            long m = (long)n - c0 < CH ? (long)n - c0 : CH;
            for (long i = 0; i < m; i++) {
                g2_set_zero(&acc[i]);
                const uint8_t *k = (const uint8_t *)(scalars + 4 * (c0 + i));
                for (int w = 0; w < 32; w++) if (k[w]) g2_add_mixed(&acc[i], &ta[w * 255 + k[w] - 1]);
            }
            g2_batch_to_affine(aff, acc, m);
            for (long i = 0; i < m; i++) {
                if (enc) g2_encode(out + 192 * (c0 + i), &aff[i], 0); else g2_store((uint64_t *)(out + 192 * (c0 + i)), &aff[i]);
            }
        }
        free(acc); free(aff);
    }
    free(ta);
}

/* =========================================================================================
 * multiexp — upstream bellman 0.1.0 multiexp.rs, restated:
 *   c = 3 if n_exponents < 32 else ceil(ln(n_exponents));
 *   one region per c-bit window (a pool task each — here an OpenMP task per window);
 *   per region: iterate exponents with their density bit; skip exp == 0; exp == 1 is added
 *   straight to the accumulator in the FIRST region only (handle_trivial); otherwise
 *   bucket[(exp >> skip) % 2^c - 1] += base; running-sum ("summation by parts") bucket reduce;
 *   regions combined high-to-low with c doublings each.
 *   Bases advance only where density is set (the query vectors hold only dense variables).
 *   A base at infinity is an error (SynthesisError::UnexpectedIdentity) -> returns -3.
 * ========================================================================================= */
static unsigned msm_window(size_t n_exp) {
    if (n_exp < 32) return 3;
    return (unsigned)ceil(log((double)n_exp));
}
static inline uint64_t repr_window(const uint64_t *e, unsigned skip, unsigned c) {
    /* (e >> skip) % 2^c on a 256-bit little-endian repr */
    unsigned w = skip / 64, b = skip % 64;
    uint64_t v = e[w] >> b;
    if (b && w + 1 < 4) v |= e[w + 1] << (64 - b);
    return v & ((1ULL << c) - 1);
}
#define DEFINE_MSM(G, AFF, JAC)                                                                        \
static int G##_multiexp(JAC *result, const AFF *bases, const uint64_t *exps, size_t n_exp,             \
                        const uint8_t *density) {                                                      \
    unsigned c = msm_window(n_exp);                                                                    \
    int n_regions = (int)((255 + c - 1) / c);   /* skip < NUM_BITS = 255 */                            \
    JAC *regions = (JAC *)malloc(sizeof(JAC) * n_regions);                                             \
    int err = 0;                                                                                       \
    _Pragma("omp parallel for schedule(dynamic, 1)")                                                   \
    for (int reg = 0; reg < n_regions; reg++) {                                                        \
        unsigned skip = reg * c;                                                                       \
        int handle_trivial = (reg == 0);                                                               \
        JAC acc; G##_set_zero(&acc);                                                                   \
        size_t nb = ((size_t)1 << c) - 1;                                                              \
        JAC *buckets = (JAC *)malloc(sizeof(JAC) * nb);                                                \
        for (size_t i = 0; i < nb; i++) G##_set_zero(&buckets[i]);                                     \
        size_t bi = 0;                                                                                 \
        for (size_t i = 0; i < n_exp; i++) {                                                           \
            if (density && !density[i]) continue;                                                      \
            const uint64_t *e = exps + 4 * i;                                                          \
            const AFF *b = &bases[bi++];                                                               \
            if ((e[0] | e[1] | e[2] | e[3]) == 0) continue;                                            \
            if (e[0] == 1 && (e[1] | e[2] | e[3]) == 0) {                                              \
                if (handle_trivial) { if (b->inf) err = -3; G##_add_mixed(&acc, b); }                  \
                continue;                                                                              \
            }                                                                                          \
            uint64_t d = repr_window(e, skip, c);                                                      \
            if (d) { if (b->inf) err = -3; G##_add_mixed(&buckets[d - 1], b); }                        \
        }                                                                                              \
        JAC running; G##_set_zero(&running);                                                           \
        for (size_t i = nb; i-- > 0;) { G##_add(&running, &buckets[i]); G##_add(&acc, &running); }     \
        free(buckets);                                                                                 \
        regions[reg] = acc;                                                                            \
    }                                                                                                  \
    JAC hi = regions[n_regions - 1];                                                                   \
    for (int reg = n_regions - 2; reg >= 0; reg--) {                                                   \
        for (unsigned k = 0; k < c; k++) G##_dbl(&hi);                                                 \
        G##_add(&hi, &regions[reg]);                                                                   \
    }                                                                                                  \
    free(regions);                                                                                     \
    *result = hi;                                                                                      \
    return err;                                                                                        \
}
DEFINE_MSM(g1, g1_aff, g1_jac)
DEFINE_MSM(g2, g2_aff, g2_jac)

/* bases in limb form (see above), scalars canonical 4xu64; out = affine limb form.
 * density (optional, one byte per exponent) as in multiexp; n_bases = popcount(density) or n. */
EXPORT int zko_g1_msm(const uint64_t *bases, const uint64_t *scalars, size_t n, const uint8_t *density, uint64_t *out) {
    size_t nb = 0;
    if (density) { for (size_t i = 0; i < n; i++) nb += density[i] != 0; } else nb = n;
    g1_aff *b = (g1_aff *)malloc(sizeof(g1_aff) * (nb ? nb : 1));
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)nb; i++) g1_load(&b[i], bases + 12 * i);
    g1_jac r; int e = g1_multiexp(&r, b, scalars, n, density);
    g1_aff a; g1_into_affine(&a, &r); g1_store(out, &a);
    free(b);
    return e;
}
EXPORT int zko_g2_msm(const uint64_t *bases, const uint64_t *scalars, size_t n, const uint8_t *density, uint64_t *out) {
    size_t nb = 0;
    if (density) { for (size_t i = 0; i < n; i++) nb += density[i] != 0; } else nb = n;
    g2_aff *b = (g2_aff *)malloc(sizeof(g2_aff) * (nb ? nb : 1));
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)nb; i++) g2_load(&b[i], bases + 24 * i);
    g2_jac r; int e = g2_multiexp(&r, b, scalars, n, density);
    g2_aff a; g2_into_affine(&a, &r); g2_store(out, &a);
    free(b);
    return e;
}

/* =========================================================================================
 * EvaluationDomain — upstream bellman 0.1.0 domain.rs, restated:
 *   omega = ROOT_OF_UNITY^(2^(S - exp)); serial_fft = bit-reversal permutation followed by
 *   log n rounds of Cooley–Tukey butterflies with w_m = omega^(n/2m); ifft scales by m^-1;
 *   coset_fft = distribute_powers(GENERATOR) then fft; icoset_fft = ifft then
 *   distribute_powers(GENERATOR^-1); divide_by_z_on_coset multiplies by (g^m - 1)^-1.
 * Data: Montgomery-form Fr, natural order in and out.
 * ========================================================================================= */
static void fr_domain_omega(fr_t *w, unsigned log_n) {
    memcpy(w->l, FR_ROOT_OF_UNITY, 32);
    for (unsigned i = log_n; i < FR_S; i++) fr_sqr(w, w);
}
static inline uint32_t bitrev32(uint32_t x, unsigned bits) {
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static void fr_fft(fr_t *a, unsigned log_n, const fr_t *omega) {
    size_t n = (size_t)1 << log_n;
    for (size_t k = 0; k < n; k++) {
        size_t rk = bitrev32((uint32_t)k, log_n);
        if (k < rk) { fr_t t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
    if (log_n == 0) return;
    /* twiddle table w^j, j < n/2 (the serial reference recomputes w incrementally; same values) */
    fr_t *tw = (fr_t *)malloc(sizeof(fr_t) * (n / 2));
    {
        int nt = 1;
#ifdef _OPENMP
        nt = omp_get_max_threads();
#endif
        size_t half = n / 2, chunk = (half + nt - 1) / nt;
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nt; t++) {
            size_t s = t * chunk, e = s + chunk < half ? s + chunk : half;
            if (s >= e) continue;
            uint64_t ex[1] = {s};
            fr_t w; fr_pow(&w, omega, ex, 1);
            for (size_t j = s; j < e; j++) { tw[j] = w; fr_mul(&w, &w, omega); }
        }
    }
    size_t m = 1;
    for (unsigned s = 0; s < log_n; s++) {
        size_t stride = n / (2 * m);
#pragma omp parallel for schedule(static)
        for (long idx = 0; idx < (long)(n / 2); idx++) {
            size_t k = ((size_t)idx / m) * 2 * m, j = (size_t)idx % m;
            fr_t t, u;
            fr_mul(&t, &a[k + j + m], &tw[j * stride]);
            u = a[k + j];
            fr_sub(&a[k + j + m], &u, &t);
            fr_add(&a[k + j], &u, &t);
        }
        m *= 2;
    }
    free(tw);
}
static void fr_distribute_powers(fr_t *a, size_t n, const fr_t *g) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    size_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        size_t s = t * chunk, e = s + chunk < n ? s + chunk : n;
        if (s >= e) continue;
        uint64_t ex[1] = {s};
        fr_t u; fr_pow(&u, g, ex, 1);
        for (size_t i = s; i < e; i++) { fr_mul(&a[i], &a[i], &u); fr_mul(&u, &u, g); }
    }
}
static void fr_scale(fr_t *a, size_t n, const fr_t *k) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) fr_mul(&a[i], &a[i], k);
}
static void fr_ifft(fr_t *a, unsigned log_n, const fr_t *omega) {
    fr_t winv, minv, m;
    fr_inv(&winv, omega);
    fr_fft(a, log_n, &winv);
    uint64_t mr[4] = {(uint64_t)1 << log_n, 0, 0, 0};
    fr_from_repr(&m, mr); fr_inv(&minv, &m);
    fr_scale(a, (size_t)1 << log_n, &minv);
}
/* mode: 0 fft, 1 ifft, 2 coset_fft, 3 icoset_fft */
EXPORT int zko_fr_ntt(uint64_t *data, unsigned log_n, int mode) {
    if (log_n > FR_S) return -2;   /* PolynomialDegreeTooLarge */
    fr_t *a = (fr_t *)data, w, g, ginv;
    size_t n = (size_t)1 << log_n;
    fr_domain_omega(&w, log_n);
    memcpy(g.l, FR_GENERATOR, 32);
    switch (mode) {
    case 0: fr_fft(a, log_n, &w); break;
    case 1: fr_ifft(a, log_n, &w); break;
    case 2: fr_distribute_powers(a, n, &g); fr_fft(a, log_n, &w); break;
    case 3: fr_ifft(a, log_n, &w); fr_inv(&ginv, &g); fr_distribute_powers(a, n, &ginv); break;
    default: return -1;
    }
    return 0;
}

/* quotient polynomial exactly as create_proof computes it; a/b/c: n_c Montgomery evaluations each
 * (padded with zeros to m = 2^log_m); out: m-1 Montgomery coefficients. */
static int h_poly(fr_t *out, const fr_t *a_in, const fr_t *b_in, const fr_t *c_in, size_t n_c, unsigned *log_m_out) {
    size_t m = 1; unsigned exp = 0;
    while (m < n_c) { m *= 2; exp++; if (exp >= FR_S) return -2; }
    fr_t *a = (fr_t *)calloc(m, sizeof(fr_t)), *b = (fr_t *)calloc(m, sizeof(fr_t)), *c = (fr_t *)calloc(m, sizeof(fr_t));
    memcpy(a, a_in, n_c * sizeof(fr_t)); memcpy(b, b_in, n_c * sizeof(fr_t)); memcpy(c, c_in, n_c * sizeof(fr_t));
    fr_t w, g, ginv, z, zinv, one;
    fr_domain_omega(&w, exp);
    memcpy(g.l, FR_GENERATOR, 32); fr_inv(&ginv, &g);
    fr_t *v[3] = {a, b, c};
    for (int k = 0; k < 3; k++) { fr_ifft(v[k], exp, &w); fr_distribute_powers(v[k], m, &g); fr_fft(v[k], exp, &w); }
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)m; i++) { fr_mul(&a[i], &a[i], &b[i]); fr_sub(&a[i], &a[i], &c[i]); }
    uint64_t me[1] = {m};
    fr_pow(&z, &g, me, 1); fr_set_one(&one); fr_sub(&z, &z, &one); fr_inv(&zinv, &z);
    fr_scale(a, m, &zinv);
    fr_ifft(a, exp, &w); fr_distribute_powers(a, m, &ginv);
    memcpy(out, a, (m - 1) * sizeof(fr_t));
    free(a); free(b); free(c);
    *log_m_out = exp;
    return 0;
}
/* canonical in / canonical out flavour for tests */
EXPORT int zko_h_coeffs(const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t n_c, uint64_t *out /* (m-1)*4 */) {
    fr_t *am = (fr_t *)malloc(n_c * 32), *bm = (fr_t *)malloc(n_c * 32), *cm = (fr_t *)malloc(n_c * 32);
    zko_fr_from_repr_many(a, (uint64_t *)am, n_c); zko_fr_from_repr_many(b, (uint64_t *)bm, n_c); zko_fr_from_repr_many(c, (uint64_t *)cm, n_c);
    size_t m = 1; while (m < n_c) m *= 2;
    fr_t *h = (fr_t *)malloc(m * 32);
    unsigned lg; int e = h_poly(h, am, bm, cm, n_c, &lg);
    if (!e) zko_fr_into_repr_many((uint64_t *)h, out, m - 1);
    free(am); free(bm); free(cm); free(h);
    return e;
}

/* =========================================================================================
 * groth16::Parameters (upstream bellman; grammar in SURVEY.md §3.3, verified against
 * zface/params/conf_pk.dat) and groth16::create_proof.
 * ========================================================================================= */
typedef struct {
    g1_aff alpha_g1, beta_g1, delta_g1; g2_aff beta_g2, gamma_g2, delta_g2;
    size_t n_ic, n_h, n_l, n_a, n_b;
    g1_aff *ic, *h, *l, *a, *b_g1; g2_aff *b_g2;
} zko_params;

static uint32_t rd_u32be(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

EXPORT void zko_params_free(zko_params *p) {
    if (!p) return;
    free(p->ic); free(p->h); free(p->l); free(p->a); free(p->b_g1); free(p->b_g2); free(p);
}
/* returns 0 or: -1 truncated/io, -10-e point decoding error e, -20 point at infinity in a query */
EXPORT int zko_params_read(const uint8_t *buf, size_t len, int checked, zko_params **out) {
    zko_params *p = (zko_params *)calloc(1, sizeof(zko_params));
    size_t off = 0; int e = 0;
#define NEED(k) do { if (off + (k) > len) { zko_params_free(p); return -1; } } while (0)
#define RD_G1(dst) do { NEED(96); e = g1_decode_uncompressed(&(dst), buf + off, checked); off += 96; if (e) { zko_params_free(p); return -10 - e; } } while (0)
#define RD_G2(dst) do { NEED(192); e = g2_decode_uncompressed(&(dst), buf + off, checked); off += 192; if (e) { zko_params_free(p); return -10 - e; } } while (0)
    RD_G1(p->alpha_g1); RD_G1(p->beta_g1); RD_G2(p->beta_g2); RD_G2(p->gamma_g2); RD_G1(p->delta_g1); RD_G2(p->delta_g2);
    size_t *cnt[5] = {&p->n_ic, &p->n_h, &p->n_l, &p->n_a, &p->n_b};
    g1_aff **vec[5] = {&p->ic, &p->h, &p->l, &p->a, &p->b_g1};
    for (int k = 0; k < 5; k++) {
        NEED(4); size_t n = rd_u32be(buf + off); off += 4;
        NEED(96 * n);
        *cnt[k] = n; *vec[k] = (g1_aff *)malloc(sizeof(g1_aff) * (n ? n : 1));
        int err = 0;
#pragma omp parallel for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; i++) {
            int ee = g1_decode_uncompressed(&(*vec[k])[i], buf + off + 96 * i, checked);
            if (ee) err = -10 - ee; else if ((*vec[k])[i].inf) err = -20;
        }
        off += 96 * n;
        if (err) { zko_params_free(p); return err; }
    }
    {
        NEED(4); size_t n = rd_u32be(buf + off); off += 4;
        NEED(192 * n);
        if (n != p->n_b) { /* b_g1 and b_g2 have independent length prefixes; keep b_g2's own */ }
        p->b_g2 = (g2_aff *)malloc(sizeof(g2_aff) * (n ? n : 1));
        int err = 0;
#pragma omp parallel for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; i++) {
            int ee = g2_decode_uncompressed(&p->b_g2[i], buf + off + 192 * i, checked);
            if (ee) err = -10 - ee; else if (p->b_g2[i].inf) err = -20;
        }
        off += 192 * n;
        if (err) { zko_params_free(p); return err; }
        if (n != p->n_b) { zko_params_free(p); return -1; }
    }
    *out = p;
    return 0;
}
EXPORT void zko_params_counts(const zko_params *p, uint64_t out[5]) { out[0] = p->n_ic; out[1] = p->n_h; out[2] = p->n_l; out[3] = p->n_a; out[4] = p->n_b; }
/* copy a query vector out in limb form: which 0 ic,1 h,2 l,3 a,4 b_g1 (96 B each), 5 b_g2 (192 B each) */
EXPORT void zko_params_export(const zko_params *p, int which, uint64_t *out) {
    const g1_aff *v[5] = {p->ic, p->h, p->l, p->a, p->b_g1};
    size_t n[5] = {p->n_ic, p->n_h, p->n_l, p->n_a, p->n_b};
    if (which < 5) for (size_t i = 0; i < n[which]; i++) g1_store(out + 12 * i, &v[which][i]);
    else for (size_t i = 0; i < p->n_b; i++) g2_store(out + 24 * i, &p->b_g2[i]);
}

/* create_proof.  All Fr inputs are canonical FrRepr limbs (as they cross the C ABI, include/zkb200.h).
 * Returns 0, or -2 PolynomialDegreeTooLarge, -3 UnexpectedIdentity, -4 AssignmentMissing/size mismatch. */
EXPORT int zko_groth16_prove(const zko_params *p,
                             const uint64_t *a_ev, const uint64_t *b_ev, const uint64_t *c_ev, size_t n_c,
                             const uint64_t *inputs, size_t n_inputs, const uint64_t *aux, size_t n_aux,
                             const uint8_t *a_aux_density, const uint8_t *b_input_density, const uint8_t *b_aux_density,
                             const uint64_t r_[4], const uint64_t s_[4], uint8_t proof_out[192]) {
    /* ---- h ---- */
    size_t m = 1; while (m < n_c) m *= 2;
    fr_t *am = (fr_t *)malloc(n_c * 32 + 32), *bm = (fr_t *)malloc(n_c * 32 + 32), *cm = (fr_t *)malloc(n_c * 32 + 32);
    zko_fr_from_repr_many(a_ev, (uint64_t *)am, n_c); zko_fr_from_repr_many(b_ev, (uint64_t *)bm, n_c); zko_fr_from_repr_many(c_ev, (uint64_t *)cm, n_c);
    fr_t *h = (fr_t *)malloc(m * 32);
    unsigned lg; int e = h_poly(h, am, bm, cm, n_c, &lg);
    free(am); free(bm); free(cm);
    if (e) { free(h); return e; }
    uint64_t *hrepr = (uint64_t *)malloc(m * 32);
    zko_fr_into_repr_many((uint64_t *)h, hrepr, m - 1);
    free(h);
    size_t a_aux_total = 0, b_in_total = 0, b_aux_total = 0;
    for (size_t i = 0; i < n_aux; i++) { a_aux_total += a_aux_density[i] != 0; b_aux_total += b_aux_density[i] != 0; }
    for (size_t i = 0; i < n_inputs; i++) b_in_total += b_input_density[i] != 0;
    /* ParameterSource bounds (get_h / get_l / get_a / get_b_*) */
    if (p->n_h < m - 1 || p->n_l < n_aux || p->n_a < n_inputs + a_aux_total || p->n_b < b_in_total + b_aux_total || p->n_ic != n_inputs) {
        free(hrepr); return -4;
    }
    g1_jac H, L, Ain, Aaux, B1in, B1aux; g2_jac B2in, B2aux;
    int err = 0;
    err |= g1_multiexp(&H, p->h, hrepr, m - 1, NULL);
    free(hrepr);
    err |= g1_multiexp(&L, p->l, aux, n_aux, NULL);
    err |= g1_multiexp(&Ain, p->a, inputs, n_inputs, NULL);
    err |= g1_multiexp(&Aaux, p->a + n_inputs, aux, n_aux, a_aux_density);
    err |= g1_multiexp(&B1in, p->b_g1, inputs, n_inputs, b_input_density);
    err |= g1_multiexp(&B1aux, p->b_g1 + b_in_total, aux, n_aux, b_aux_density);
    err |= g2_multiexp(&B2in, p->b_g2, inputs, n_inputs, b_input_density);
    err |= g2_multiexp(&B2aux, p->b_g2 + b_in_total, aux, n_aux, b_aux_density);
    if (err) return -3;
    if (p->delta_g1.inf || p->delta_g2.inf) return -3;
    fr_t r, s, rs; uint64_t rs_repr[4];
    if (fr_from_repr(&r, r_) || fr_from_repr(&s, s_)) return -4;
    fr_mul(&rs, &r, &s); fr_into_repr(rs_repr, &rs);
    g1_jac g_a, g_c, t; g2_jac g_b;
    g1_aff_mul(&g_a, &p->delta_g1, r_); g1_add_mixed(&g_a, &p->alpha_g1);
    g2_aff_mul(&g_b, &p->delta_g2, s_); g2_add_mixed(&g_b, &p->beta_g2);
    g1_aff_mul(&g_c, &p->delta_g1, rs_repr);
    g1_aff_mul(&t, &p->alpha_g1, s_); g1_add(&g_c, &t);
    g1_aff_mul(&t, &p->beta_g1, r_); g1_add(&g_c, &t);
    g1_jac a_ans = Ain; g1_add(&a_ans, &Aaux);
    g1_add(&g_a, &a_ans);
    g1_mul(&t, &a_ans, s_); g1_add(&g_c, &t);
    g1_jac b1 = B1in; g1_add(&b1, &B1aux);
    g2_jac b2 = B2in; g2_add(&b2, &B2aux);
    g2_add(&g_b, &b2);
    g1_mul(&t, &b1, r_); g1_add(&g_c, &t);
    g1_add(&g_c, &H); g1_add(&g_c, &L);
    g1_aff pa, pc; g2_aff pb;
    g1_into_affine(&pa, &g_a); g2_into_affine(&pb, &g_b); g1_into_affine(&pc, &g_c);
    /* Proof::write (core/bellman-verifier/src/lib.rs:55-65) */
    g1_encode(proof_out, &pa, 1); g2_encode(proof_out + 48, &pb, 1); g1_encode(proof_out + 144, &pc, 1);
    return 0;
}

/* ---- verifier side ---- */
#include "pairing_oracle.inc"

EXPORT int zko_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
EXPORT void zko_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
