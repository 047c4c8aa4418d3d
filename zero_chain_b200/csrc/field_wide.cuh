// Separated multiply / Montgomery-reduce for Fq (experimental; round-1 microbenchmark material).
//
// field.cuh's operator* interleaves product and reduction rows (CIOS).  Splitting them makes three savings
// possible in the point formulas: (1) squarings need 78 instead of 144 limb products, (2) Karatsuba on the
// 12-limb product (3 x 6x6 + one 7x7 correction: ~121 products), (3) sums/differences of products can share
// ONE reduction (e.g. Y3 = R (Q - X3) - Y1 PPP in the mixed addition).  Everything here is column-wise
// ("Comba"): a three-word accumulator (c0, c1, c2) takes each 32x32 product as mad.lo.cc / madc.hi.cc / addc.
// Results are bit-identical to operator* (same Montgomery radix, fully reduced outputs).
#pragma once
#include "field.cuh"

struct FqWide { uint32_t l[24]; };

namespace zkwide {
using namespace zkprim;

// (c0,c1,c2) += a * b
ZK_DEV void mac(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t a, uint32_t b) {
    c0 = mad_lo_cc(a, b, c0);
    c1 = madc_hi_cc(a, b, c1);
    c2 = addc(c2, 0);
}
// (c0,c1,c2) += v
ZK_DEV void acc_add(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t v) {
    c0 = add_cc(c0, v);
    c1 = addc_cc(c1, 0);
    c2 = addc(c2, 0);
}

// T = a * b (24 limbs), schoolbook by columns
ZK_DEV FqWide mul_wide(const Fq &a, const Fq &b) {
    FqWide t;
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int k = 0; k < 23; k++) {
#pragma unroll
        for (int i = 0; i < 12; i++) {
            int j = k - i;
            if (j >= 0 && j < 12) mac(c0, c1, c2, a.l[i], b.l[j]);
        }
        t.l[k] = c0; c0 = c1; c1 = c2; c2 = 0;
    }
    t.l[23] = c0;
    return t;
}
// T = a^2: cross products once, doubled, plus the diagonal
ZK_DEV FqWide sqr_wide(const Fq &a) {
    FqWide t;
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int k = 0; k < 23; k++) {
        // cross terms i < j, i + j = k
        uint32_t x0 = 0, x1 = 0, x2 = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            int j = k - i;
            if (j > i && j < 12) mac(x0, x1, x2, a.l[i], a.l[j]);
        }
        // double the cross sum (3 words, cannot overflow: at most 6 products per column)
        x2 = (x2 << 1) | (x1 >> 31); x1 = (x1 << 1) | (x0 >> 31); x0 <<= 1;
        c0 = add_cc(c0, x0); c1 = addc_cc(c1, x1); c2 = addc(c2, x2);
        if ((k & 1) == 0) mac(c0, c1, c2, a.l[k >> 1], a.l[k >> 1]);
        t.l[k] = c0; c0 = c1; c1 = c2; c2 = 0;
    }
    t.l[23] = c0;
    return t;
}
// Montgomery reduction of T < p * 2^382 (so the result before the final subtraction is < 2p): T * 2^-384 mod p
ZK_DEV Fq redc(const FqWide &t) {
    uint32_t m[12];
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
#pragma unroll
        for (int j = 0; j < i; j++) mac(c0, c1, c2, m[j], FqParams::modc(i - j));
        acc_add(c0, c1, c2, t.l[i]);
        m[i] = c0 * FqParams::INV;
        mac(c0, c1, c2, m[i], FqParams::modc(0));      // c0 becomes 0
        c0 = c1; c1 = c2; c2 = 0;
    }
    Fq r;
#pragma unroll
    for (int i = 12; i < 24; i++) {
#pragma unroll
        for (int j = i - 11; j < 12; j++) mac(c0, c1, c2, m[j], FqParams::modc(i - j));
        acc_add(c0, c1, c2, t.l[i]);
        r.l[i - 12] = c0; c0 = c1; c1 = c2; c2 = 0;
    }
    return Fq::reduce_once(r);      // c0 (carry out) is zero under the input bound
}
// a - b + p * 2^381 (keeps the difference of two products non-negative and below p * 2^382)
ZK_DEV FqWide sub_offset(const FqWide &a, const FqWide &b) {
    // p << 381 = p << (11*32 + 29): limbs 11..23
    FqWide off;
#pragma unroll
    for (int i = 0; i < 24; i++) off.l[i] = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint32_t lo = FqParams::mod(i) << 29, hi = FqParams::mod(i) >> 3;
        off.l[11 + i] |= lo;
        off.l[12 + i] |= hi;
    }
    FqWide r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 24; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
    // borrow is absorbed by the offset (two's complement wrap is fine: the true value + offset fits in 24 limbs)
    r.l[0] = add_cc(r.l[0], off.l[0]);
#pragma unroll
    for (int i = 1; i < 23; i++) r.l[i] = addc_cc(r.l[i], off.l[i]);
    r.l[23] = addc(r.l[23], off.l[23]);
    return r;
}
ZK_DEV Fq mul_sep(const Fq &a, const Fq &b) { return redc(mul_wide(a, b)); }
ZK_DEV Fq sqr_sep(const Fq &a) { return redc(sqr_wide(a)); }
// a*b - c*d with one reduction
ZK_DEV Fq mul_sub_mul(const Fq &a, const Fq &b, const Fq &c, const Fq &d) { return redc(sub_offset(mul_wide(a, b), mul_wide(c, d))); }

}  // namespace zkwide
