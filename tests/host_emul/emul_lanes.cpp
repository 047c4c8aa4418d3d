// CPU unit-test harness for the WARP-COOPERATIVE device code: zero_chain_b200/csrc/pairing_lanes.cuh (an Fq12 value spread
// over six lanes) and curve_coop.cuh (one point operation spread over the lanes of a warp).  Same source as the device build;
// the limb arithmetic comes from ZK_HOST_EMUL (field.cuh), and the SIMT pieces the headers use are emulated here: every lane is
// a host thread, `__shfl_sync` / `__ballot_sync` exchange values through a slot array and a spinning barrier (the device code
// keeps its control flow warp-uniform, so all lanes reach the same exchanges in the same order).
// Test infrastructure only — never linked into libzkb200.so.
#define ZK_HOST_EMUL 1
#include <atomic>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>

// ---- SIMT shim ---------------------------------------------------------------------------------------------------------
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __constant__
struct EmuDim { unsigned x; };
static thread_local EmuDim threadIdx;
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 __ldg(const uint4 *p) { return *p; }

namespace simt {
static int n_lanes = 0;
static std::atomic<int> arrived{0};
static std::atomic<unsigned> phase{0};
static uint32_t slots[2][32];
static thread_local unsigned turn = 0;
static void barrier() {
    const unsigned ph = phase.load(std::memory_order_acquire);
    if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == n_lanes) {
        arrived.store(0, std::memory_order_relaxed);
        phase.store(ph + 1, std::memory_order_release);
    } else {
        while (phase.load(std::memory_order_acquire) == ph) std::this_thread::yield();
    }
}
// one barrier per exchange: the slot buffers alternate, and nobody can start exchange k + 2 before everybody has left k
static uint32_t exchange(uint32_t v, int src) {
    const int lane = threadIdx.x & 31;
    uint32_t *s = slots[turn & 1];
    turn++;
    s[lane] = v;
    barrier();
    return s[src < n_lanes ? src : lane];
}
template <class Fn>
static void run(int lanes, Fn fn) {
    n_lanes = lanes; arrived = 0; phase = 0;
    std::vector<std::thread> th;
    for (int l = 0; l < lanes; l++) th.emplace_back([=] { threadIdx.x = (unsigned)l; turn = 0; fn(l); });
    for (auto &t : th) t.join();
}
}  // namespace simt
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return simt::exchange(v, src); }
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, int delta) { const int l = threadIdx.x & 31; return simt::exchange(v, l - delta >= 0 ? l - delta : l); }
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, int delta) { const int l = threadIdx.x & 31; return simt::exchange(v, l + delta < simt::n_lanes ? l + delta : l); }
static inline unsigned __ballot_sync(unsigned, bool v) {
    unsigned m = 0;
    for (int l = 0; l < simt::n_lanes; l++) m |= (simt::exchange(v ? 1u : 0u, l) & 1u) << l;
    return m;
}

#include "pairing_lanes.cuh"
#include "curve_coop.cuh"
#include "msm_warp_scan.cuh"
using namespace zkpair;
using namespace zklanes;

// ---- Fq12 on six lanes ---------------------------------------------------------------------------------------------------
// a, b, out: Fq12 in tower memory order (576 bytes).  op: 0 mul12, 1 sqr12, 2 cyclotomic_sqr, 3 inv12, 4 conj12,
// 5..7 frobenius12 k = 1..3, 8 final_exponentiation, 9 mul_w2, 10 exp_x.  lanes = 6 (one group) or 32 (five groups + the two
// shadow lanes; every group computes the same thing and must agree: the return value counts disagreeing groups).
extern "C" int emu_lanes_f12(int op, int lanes, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    const Fq2 *A = reinterpret_cast<const Fq2 *>(a), *B = reinterpret_cast<const Fq2 *>(b);
    std::vector<Fq2> res(32);
    simt::run(lanes, [&](int lane) {
        const Lane L = Lane::make();
        const Fq2 x = A[slot_index(L.t)], y = b ? B[slot_index(L.t)] : Fq2::zero();
        Fq2 r;
        switch (op) {
            case 0: r = mul12(L, x, y); break;
            case 1: r = sqr12(L, x); break;
            case 2: r = cyclotomic_sqr(L, x); break;
            case 3: r = inv12(L, x); break;
            case 4: r = conj12(L, x); break;
            case 5: case 6: case 7: { const FrobConsts fc = frob_consts(L); r = frobenius12(L, fc, x, op - 4); break; }
            case 8: r = final_exponentiation(L, x); break;
            case 9: r = mul_w2(L, x); break;
            default: r = exp_x(L, x); break;
        }
        res[lane] = r;
    });
    Fq2 *O = reinterpret_cast<Fq2 *>(out);
    for (int t = 0; t < 6; t++) O[slot_index(t)] = res[t];
    int bad = 0;
    for (int l = 6; l < lanes; l++) {
        const int t = l < 30 ? l % 6 : l - 30;
        if (memcmp(&res[l], &res[t], sizeof(Fq2)) != 0) bad++;
    }
    return bad;
}
// AND over the six lanes of a group (all_lanes): lane `off` of every group votes false, everybody else true
extern "C" int emu_lanes_all(int lanes, int off) {
    std::vector<int> res(32);
    simt::run(lanes, [&](int lane) {
        const Lane L = Lane::make();
        res[lane] = all_lanes(L, off < 0 || L.t != off) ? 1 : 0;
    });
    int s = 0;
    for (int l = 0; l < lanes; l++) s += res[l];
    return s;
}
// merged Miller loop of three pairs: p = 3 affine G1 points (Montgomery limbs, 96 B each), coeffs = 3 x N_COEFFS prepared line
// coefficients (g2_prepare), skip bit k = pair k has a point at infinity
extern "C" void emu_lanes_miller3(const uint32_t *p, const uint32_t *coeffs, int skip, uint32_t *out) {
    Fq pxy[3][2];
    memcpy(pxy, p, sizeof(pxy));
    const LineCoeff *c = reinterpret_cast<const LineCoeff *>(coeffs);
    std::vector<Fq2> res(6);
    simt::run(6, [&](int lane) {
        const Lane L = Lane::make();
        PairIn in[3];
        for (int k = 0; k < 3; k++) { in[k].pxy = pxy[k]; in[k].coeffs = c + (size_t)k * N_COEFFS; in[k].stride = 1; in[k].skip = (skip >> k) & 1; }
        res[lane] = miller_loop3(L, in[0], in[1], in[2]);
    });
    Fq2 *O = reinterpret_cast<Fq2 *>(out);
    for (int t = 0; t < 6; t++) O[slot_index(t)] = res[t];
}
extern "C" int emu_n_coeffs() { return N_COEFFS; }

// ---- cooperative XYZZ operations (curve_coop.cuh) against the one-thread formulas of curve.cuh ------------------------------
// a, b: affine points; the operands are first made projective with non-trivial ZZ (2a, 3b) so that every product of the formulas
// matters.  op 0: dbl(2a), 1: 2a + 3b, 2: 2a + 2a (doubling inside add), 3: 2a + (-2a), 4: inf + 3b, 5: 2a + inf.
// out = the affine result of the cooperative path; returns the number of lanes whose result differs from the one-thread formulas.
template <class F>
static int t_coop(int op, int lanes, const uint32_t *pa, const uint32_t *pb, uint32_t *out) {
    Affine<F> a, b;
    memcpy(&a, pa, sizeof(a)); memcpy(&b, pb, sizeof(b));
    XYZZ<F> P = XYZZ<F>::from_affine(a).dbl(), Q = XYZZ<F>::from_affine(b).dbl();
    Q.add_mixed(b);
    XYZZ<F> want = P;
    switch (op) {
        case 0: want = P.dbl(); break;
        case 1: want.add(Q); break;
        case 2: Q = P; want.add(Q); break;
        case 3: Q = P; Q.y = Q.y.neg(); want.add(Q); break;
        case 4: P = XYZZ<F>::inf(); want = Q; break;
        default: Q = XYZZ<F>::inf(); break;
    }
    std::vector<XYZZ<F>> res(lanes);
    simt::run(lanes, [&](int lane) {
        XYZZ<F> x = P;
        if (op == 0) zkcoop::dbl(x); else zkcoop::add(x, Q);
        res[lane] = x;
    });
    const Affine<F> w = want.to_affine();
    int bad = 0;
    for (int l = 0; l < lanes; l++) {
        const Affine<F> g = res[l].to_affine();
        if (memcmp(&g, &w, sizeof(w)) != 0) bad++;
    }
    const Affine<F> r = res[0].to_affine();
    memcpy(out, &r, sizeof(r));
    return bad;
}
extern "C" int emu_coop_g1(int op, int lanes, const uint32_t *a, const uint32_t *b, uint32_t *o) { return t_coop<Fq>(op, lanes, a, b, o); }
extern "C" int emu_coop_g2(int op, int lanes, const uint32_t *a, const uint32_t *b, uint32_t *o) { return t_coop<Fq2>(op, lanes, a, b, o); }

// ---- warp products of the batched-affine rounds (msm_warp_scan.cuh) ---------------------------------------------------------
// totals: 32 field elements (one per lane); others[l] = product of the other 31, all[l] = product of all 32 (every lane)
template <class F>
static void t_warp_products(const uint32_t *totals, uint32_t *others, uint32_t *all) {
    const F *T = reinterpret_cast<const F *>(totals);
    F *O = reinterpret_cast<F *>(others), *A = reinterpret_cast<F *>(all);
    simt::run(32, [&](int lane) { zkmsm::ba_warp_products(T[lane], O[lane], A[lane]); });
}
extern "C" void emu_warp_products_fq(const uint32_t *t, uint32_t *o, uint32_t *a) { t_warp_products<Fq>(t, o, a); }
extern "C" void emu_warp_products_fq2(const uint32_t *t, uint32_t *o, uint32_t *a) { t_warp_products<Fq2>(t, o, a); }
