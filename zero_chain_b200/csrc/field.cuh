// BLS12-381 prime-field arithmetic for sm_100a: Fq (12 x u32) and Fr (8 x u32), Montgomery form.
//
// Replaces, bit-for-bit, the reference's 64-bit-limb CPU arithmetic:
//   Fq  core/pairing/src/bls12_381/fq.rs   mul_assign 915-965 + mont_reduce 1042-1127,
//       add/sub/double/negate 818-852, from_repr 752-761, into_repr 764-773, inverse 854-907
//   Fr  core/pairing/src/bls12_381/fr.rs   mul_assign 438-464 + mont_reduce 520-571, 341-376
// Same Montgomery radix (R = 2^384 / 2^256) and the same invariant — every value is fully reduced
// (< modulus) — so a limb dump of any intermediate equals the reference's limbs.
//
// Multiplication is an interleaved (CIOS) Montgomery product on 32-bit limbs with the even/odd
// accumulator split: products a[j]*b_i for even j and odd j go to two accumulators whose 64-bit
// partial products never overlap, so each row is one unbroken mad.lo.cc / madc.hi.cc carry chain
// (ptxas fuses each lo/hi pair into IMAD.WIDE.U32 with carry-in/out).  After each row's reduction
// the accumulators swap roles, which performs the division by 2^32 without moving data.
//
// The arithmetic primitives have a host emulation (ZK_HOST_EMUL) used ONLY by the CPU unit test
// tests/host_emul (same C++ source, explicit carry flag) so the algorithm can be validated where
// there is no GPU.  The product library never compiles that path.
#pragma once
#include <stdint.h>

// Inlining policy.  A Montgomery product is ~770 SASS instructions; inlining it at every call site
// of the cold paths (G2 arithmetic, scalar multiplication, decoding, reductions) makes ptxas run for
// tens of minutes.  Translation units that hold the hot kernels define ZK_HOT and get everything
// inlined; all others call the product and the point operations as real functions.
#ifdef ZK_HOST_EMUL
#define ZK_DEV inline
#define ZK_MULFN inline
#define ZK_PTFN inline
#define ZK_F2FN inline
#elif defined(ZK_HOT)
#define ZK_MULFN __device__ __forceinline__
#define ZK_PTFN __device__ __forceinline__
#elif defined(ZK_SEMI_HOT)
// pairing.cu: the Fq product is inlined into its callers, and the Fq2 product / square become the call boundary, so the
// three (two) independent Montgomery chains of one Fq2 operation interleave in the integer pipe
#define ZK_MULFN __device__ __forceinline__
#define ZK_PTFN __device__ __noinline__
#define ZK_F2FN __device__ __noinline__
#else
#define ZK_MULFN __device__ __noinline__
#define ZK_PTFN __device__ __noinline__
#endif
#ifdef ZK_HOST_EMUL
namespace zkprim {
static thread_local uint32_t cf = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cf; cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cf; cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + cf; }
}  // namespace zkprim
#else
#define ZK_DEV __device__ __forceinline__
#ifndef ZK_F2FN
#define ZK_F2FN ZK_DEV
#endif
namespace zkprim {
// The PTX condition-code register carries between consecutive asm volatile statements (the
// established CGBN / sppark idiom): volatile asms are not reordered against each other.
ZK_DEV uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_DEV uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_DEV uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_DEV uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_DEV uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
}  // namespace zkprim
#endif

// ---- field parameter packs (fq.rs:5-43, fr.rs:4-55; 32-bit little-endian limbs) ----------------
// The modulus is also kept in the constant bank: as an immediate ptxas cannot fuse the reduction's
// mad.lo.cc/madc.hi.cc pairs into IMAD.WIDE.U32.X, as a c[bank][off] operand it can (and it costs
// no registers).
#ifndef ZK_HOST_EMUL
static __device__ __constant__ uint32_t ZK_FQ_MOD[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                                         0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
static __device__ __constant__ uint32_t ZK_FR_MOD[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
#endif
struct FqParams {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;   // -q^-1 mod 2^32 (low word of fq.rs:43)
#ifndef ZK_HOST_EMUL
    ZK_DEV static uint32_t modc(int i) { return ZK_FQ_MOD[i]; }
#else
    ZK_DEV static uint32_t modc(int i) { return mod(i); }
#endif
    ZK_DEV static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return m[i];
    }
    ZK_DEV static constexpr uint32_t one(int i) {   // R = 2^384 mod q
        constexpr uint32_t m[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return m[i];
    }
    ZK_DEV static constexpr uint32_t r2(int i) {    // R^2 mod q
        constexpr uint32_t m[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return m[i];
    }
};
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t INV = 0xffffffffu;   // -r^-1 mod 2^32 (low word of fr.rs:36)
#ifndef ZK_HOST_EMUL
    ZK_DEV static uint32_t modc(int i) { return ZK_FR_MOD[i]; }
#else
    ZK_DEV static uint32_t modc(int i) { return mod(i); }
#endif
    ZK_DEV static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return m[i];
    }
    ZK_DEV static constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return m[i];
    }
    ZK_DEV static constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return m[i];
    }
};

template <class P>
struct alignas(16) Fp {
    static constexpr int N = P::N;
    uint32_t l[N];

    ZK_DEV static Fp zero() { Fp r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
    ZK_DEV static Fp one() { Fp r; for (int i = 0; i < N; i++) r.l[i] = P::one(i); return r; }
    ZK_DEV bool is_zero() const { uint32_t o = 0; for (int i = 0; i < N; i++) o |= l[i]; return o == 0; }
    ZK_DEV bool operator==(const Fp &b) const { uint32_t o = 0; for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i]; return o == 0; }
    ZK_DEV bool operator!=(const Fp &b) const { return !(*this == b); }

    // r = (t >= p) ? t - p : t      (fq.rs:1028-1036 `reduce`)
    ZK_DEV static Fp reduce_once(const Fp &t) {
        using namespace zkprim;
        Fp u;
        u.l[0] = sub_cc(t.l[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) u.l[i] = subc_cc(t.l[i], P::mod(i));
        uint32_t borrow = subc(0, 0);   // 0xffffffff when t < p
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? t.l[i] : u.l[i];
        return r;
    }
    ZK_DEV friend Fp operator+(const Fp &a, const Fp &b) {
        using namespace zkprim;
        Fp t;
        t.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) t.l[i] = addc_cc(a.l[i], b.l[i]);
        t.l[N - 1] = addc(a.l[N - 1], b.l[N - 1]);      // moduli leave spare top bits: no carry out
        return reduce_once(t);
    }
    ZK_DEV friend Fp operator-(const Fp &a, const Fp &b) {
        using namespace zkprim;
        Fp t;
        t.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) t.l[i] = subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = subc(0, 0);
        Fp r;
        r.l[0] = add_cc(t.l[0], borrow & P::mod(0));
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(t.l[i], borrow & P::mod(i));
        r.l[N - 1] = addc(t.l[N - 1], borrow & P::mod(N - 1));
        return r;
    }
    ZK_DEV Fp dbl() const { return *this + *this; }
    ZK_DEV Fp neg() const { return is_zero() ? *this : sub_raw_mod(*this); }
    ZK_DEV static Fp sub_raw_mod(const Fp &a) {   // p - a, a != 0
        using namespace zkprim;
        Fp r;
        r.l[0] = sub_cc(P::mod(0), a.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(P::mod(i), a.l[i]);
        r.l[N - 1] = subc(P::mod(N - 1), a.l[N - 1]);
        return r;
    }
    ZK_DEV Fp cneg(bool flag) const { return flag ? neg() : *this; }

    // ---- Montgomery product ------------------------------------------------------------------
    // acc[j], acc[j+1] (j even) += a[j] * bi, one carry chain; returns nothing, leaves carry in CC.
    ZK_DEV static void row_first(uint32_t *even, uint32_t *odd, const uint32_t *a, uint32_t bi) {
        using namespace zkprim;
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            even[j] = mul_lo(a[j], bi); even[j + 1] = mul_hi(a[j], bi);
            odd[j] = mul_lo(a[j + 1], bi); odd[j + 1] = mul_hi(a[j + 1], bi);
        }
    }
    template <class A>
    ZK_DEV static void cmad_even(uint32_t *acc, A a, uint32_t bi) {   // acc += sum_{j even} a(j) bi 2^(32j); carry left in CC
        using namespace zkprim;
        acc[0] = mad_lo_cc(a(0), bi, acc[0]);
        acc[1] = madc_hi_cc(a(0), bi, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = madc_lo_cc(a(j), bi, acc[j]);
            acc[j + 1] = madc_hi_cc(a(j), bi, acc[j + 1]);
        }
    }
    // One CIOS row: (even, odd) <- ((even, odd) + a*bi + m*p) / 2^32 with roles swapped on return:
    // on entry `even[k]` sits at limb k and `odd[k]` at limb k+1 of the running total, but `odd`
    // still holds the previous row's even array (its limb 0 is zero and limb 1 is pending).
    ZK_DEV static void row(uint32_t *even, uint32_t *odd, const uint32_t *a, uint32_t bi) {
        using namespace zkprim;
        // fold the pending limb and shift the old-even array down by two limbs while adding odd products
        even[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {
            odd[j] = madc_lo_cc(a[j + 1], bi, odd[j + 2]);
            odd[j + 1] = madc_hi_cc(a[j + 1], bi, odd[j + 3]);
        }
        odd[N - 2] = madc_lo_cc(a[N - 1], bi, 0);
        odd[N - 1] = madc_hi(a[N - 1], bi, 0);
        cmad_even(even, [&](int j) { return a[j]; }, bi);
        odd[N - 1] = addc(odd[N - 1], 0);
        reduce_row(even, odd);
    }
    ZK_DEV static void reduce_row(uint32_t *even, uint32_t *odd) {
        using namespace zkprim;
        uint32_t mi = even[0] * P::INV;
        // odd += sum_{j odd} p[j] mi 2^(32(j-1))   (no carry out: total stays < 2^(32(N+1)))
        odd[0] = mad_lo_cc(P::modc(1), mi, odd[0]);
        odd[1] = madc_hi_cc(P::modc(1), mi, odd[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            odd[j] = madc_lo_cc(P::modc(j + 1), mi, odd[j]);
            odd[j + 1] = madc_hi_cc(P::modc(j + 1), mi, odd[j + 1]);
        }
        cmad_even(even, [](int j) { return P::modc(j); }, mi);
        odd[N - 1] = addc(odd[N - 1], 0);
    }
    ZK_MULFN friend Fp operator*(const Fp &a, const Fp &b) {
        using namespace zkprim;
        uint32_t even[N], odd[N];
        row_first(even, odd, a.l, b.l[0]);
        reduce_row(even, odd);
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            row(odd, even, a.l, b.l[i]);
            if (i + 1 < N) row(even, odd, a.l, b.l[i + 1]);
        }
        // N is even, so the last call was row(odd, even, ..): `odd` was just reduced (odd[0] == 0, its
        // limbs 1.. are pending one position down) and `even` sits at limb 0 of the quotient.
        Fp t;
        t.l[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) t.l[k] = addc_cc(even[k], odd[k + 1]);
        t.l[N - 1] = addc(even[N - 1], 0);
        return reduce_once(t);
    }
    ZK_DEV Fp sqr() const { return *this * *this; }

    // canonical (non-Montgomery) integer -> Montgomery: x * R2 * R^-1   (from_repr, fq.rs:752-761)
    ZK_DEV static Fp from_canonical(const Fp &x) { Fp r2; for (int i = 0; i < N; i++) r2.l[i] = P::r2(i); return x * r2; }
    // Montgomery -> canonical: x * 1 * R^-1   (into_repr, fq.rs:764-773)
    ZK_DEV Fp to_canonical() const { Fp o = zero(); o.l[0] = 1; return *this * o; }
    ZK_DEV static bool canonical_lt_mod(const Fp &x) {   // x < p ?
        using namespace zkprim;
        sub_cc(x.l[0], P::mod(0));
        uint32_t t;
#pragma unroll
        for (int i = 1; i < N; i++) t = subc_cc(x.l[i], P::mod(i));
        (void)t;
        return subc(0, 0) != 0;
    }
    // a^e for a public exponent given as little-endian u32 words (MSB-first square-and-multiply)
    ZK_PTFN Fp pow(const uint32_t *e, int nwords) const {
        Fp acc = one();
        bool started = false;
        for (int i = nwords * 32 - 1; i >= 0; i--) {
            if (started) acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) { acc = acc * *this; started = true; }
        }
        return acc;
    }
    // raw-limb helpers for the inversion
    ZK_DEV static void raw_shr1(uint32_t *x) {
#pragma unroll
        for (int i = 0; i < N - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
        x[N - 1] >>= 1;
    }
    ZK_DEV static void raw_half_mod(uint32_t *x) {   // x <- x/2 mod p for x < p: (x even ? x : x + p) >> 1
        using namespace zkprim;
        uint32_t m = (x[0] & 1u) ? 0xffffffffu : 0u;
        x[0] = add_cc(x[0], m & P::mod(0));
#pragma unroll
        for (int i = 1; i < N - 1; i++) x[i] = addc_cc(x[i], m & P::mod(i));
        x[N - 1] = addc(x[N - 1], m & P::mod(N - 1));      // no carry out: p has spare top bits
        raw_shr1(x);
    }
    ZK_DEV static bool raw_is_one(const uint32_t *x) {
        uint32_t o = x[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; i++) o |= x[i];
        return o == 0;
    }
    // x -= y, returns true when the subtraction borrowed (x < y); x is left modified only by the caller's choice
    ZK_DEV static bool raw_lt(const uint32_t *x, const uint32_t *y) {
        using namespace zkprim;
        sub_cc(x[0], y[0]);
        uint32_t t = 0;
#pragma unroll
        for (int i = 1; i < N; i++) t = subc_cc(x[i], y[i]);
        (void)t;
        return subc(0, 0) != 0;
    }
    ZK_DEV static void raw_sub(uint32_t *x, const uint32_t *y) {
        using namespace zkprim;
        x[0] = sub_cc(x[0], y[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) x[i] = subc_cc(x[i], y[i]);
        x[N - 1] = subc(x[N - 1], y[N - 1]);
    }
    // a^-1 by the binary extended Euclid of the reference (fq.rs:854-907 / fr.rs:378-431): starting from
    // (u, b) = (a, R^2), (v, c) = (p, 0) it keeps u*R^2 = b*a and v*R^2 = c*a (mod p) and returns b or c when
    // u or v reaches 1 — the Montgomery form of the inverse.  ~20x fewer instructions than Fermat's a^(p-2),
    // which matters because a single GPU thread runs this serially.  Returns zero for zero (callers test
    // is_zero first, matching `inverse()` returning None).
    ZK_PTFN Fp inverse() const {
        if (is_zero()) return *this;
        uint32_t u[N], v[N];
        Fp b, c = zero();
#pragma unroll
        for (int i = 0; i < N; i++) { u[i] = l[i]; v[i] = P::mod(i); b.l[i] = P::r2(i); }
        while (!raw_is_one(u) && !raw_is_one(v)) {
            while (!(u[0] & 1u)) { raw_shr1(u); raw_half_mod(b.l); }
            while (!(v[0] & 1u)) { raw_shr1(v); raw_half_mod(c.l); }
            if (raw_lt(v, u)) { raw_sub(u, v); b = b - c; }
            else { raw_sub(v, u); c = c - b; }
        }
        return raw_is_one(u) ? b : c;
    }
    // Fermat inverse a^(p-2); kept as an independent cross-check of inverse() in the tests.
    ZK_PTFN Fp inverse_fermat() const {
        uint32_t e[N];
        uint32_t borrow = 2;
        for (int i = 0; i < N; i++) { uint32_t v = P::mod(i); e[i] = v - borrow; borrow = (v < borrow) ? 1u : 0u; }
        return pow(e, N);
    }
};

typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;
