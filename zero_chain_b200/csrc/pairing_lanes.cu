// Lane-parallel verifier kernels (pairing_lanes.cuh): six lanes per Fq12 value, five values per warp.  Hot translation unit:
// every field product is inlined, nothing is called and nothing spills.
#define ZK_HOT 1
#include "internal.h"
#include "pairing_lanes.cuh"

using namespace zklanes;
typedef Affine<Fq> G1A;
constexpr uint8_t zkcodec_DEC_INFINITY = 6;      // zkcodec::DEC_INFINITY (codec.cuh; not included here to keep this unit small)

// one group of six lanes per proof: f[i] = conj of the product of the three Miller loops (A_i, B_i), (acc_i, -gamma), (C_i, -delta)
// 255 registers (2 CTAs / SM): a 168-register build spilled inside the Fq2 product and measured slower (profiles/r02_experiments.md §3)
static __global__ void __launch_bounds__(128, 2) k_miller_lanes(size_t n, const G1A *__restrict__ a, const G1A *__restrict__ acc, const G1A *__restrict__ c,
                                                                const zkpair::LineCoeff *__restrict__ coef_b, const zkpair::LineCoeff *__restrict__ gamma, int gamma_inf,
                                                                const zkpair::LineCoeff *__restrict__ delta, int delta_inf, const uint8_t *__restrict__ st,
                                                                zkpair::Fq12 *__restrict__ f) {
    const Lane L = Lane::make();
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t i = warp * GROUPS_PER_WARP + L.base / 6;
    const bool in_range = i < n;
    if (!in_range) i = n - 1;                              // shadow work: the shuffles need every lane of the warp
    const bool rejected = (st[3 * i] | st[3 * i + 1] | st[3 * i + 2]) != 0;     // rejected by Proof::read: no pairing is computed
    // the three G1 points of the proof, shared by the six lanes of the group: lanes 0..5 each bring one coordinate into shared memory
    __shared__ Fq pts[4][GROUPS_PER_WARP][6];
    Fq *mine = pts[(threadIdx.x >> 5) & 3][L.base / 6];
    {
        const G1A *src = L.t < 2 ? a + i : (L.t < 4 ? acc + i : c + i);
        if (L.live) mine[L.t] = ldg_fq(reinterpret_cast<const Fq *>(src) + (L.t & 1));
    }
    __syncwarp();
    const bool inf0 = mine[0].is_zero() && mine[1].is_zero(), inf1 = mine[2].is_zero() && mine[3].is_zero(), inf2 = mine[4].is_zero() && mine[5].is_zero();
    PairIn p0{mine, coef_b + i, n, inf0};                                     // infinity on either side contributes 1 (mod.rs:50-54)
    PairIn p1{mine + 2, gamma, 1, inf1 || gamma_inf != 0};
    PairIn p2{mine + 4, delta, 1, inf2 || delta_inf != 0};
    const Fq2 r = miller_loop3(L, p0, p1, p2);
    if (in_range && L.live && !rejected) reinterpret_cast<Fq2 *>(f + i)[slot_index(L.t)] = r;
}
// verdict: 1 = Ok(true), 0 = Ok(false), 2 = Proof::read -> InvalidData, 3 = Proof::read -> PointInfinity (first failing point)
static __global__ void __launch_bounds__(128, 2) k_verify_final_lanes(size_t n, const zkpair::Fq12 *__restrict__ f, const zkpair::Fq12 *__restrict__ alpha_beta,
                                                                      const uint8_t *__restrict__ st, uint8_t *__restrict__ verdict) {
    const Lane L = Lane::make();
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t i = warp * GROUPS_PER_WARP + L.base / 6;
    const bool in_range = i < n;
    if (!in_range) i = n - 1;
    uint8_t bad = 0;
    for (int s = 2; s >= 0; s--) { uint8_t e = st[3 * i + s]; if (e) bad = e == zkcodec_DEC_INFINITY ? 3 : 2; }
    Fq2 m = bad ? (L.t == 0 ? Fq2::one() : Fq2::zero()) : reinterpret_cast<const Fq2 *>(f + i)[slot_index(L.t)];
    const bool zero = all_lanes(L, m.is_zero());           // Engine::final_exponentiation -> None for f = 0
    if (zero) m = L.t == 0 ? Fq2::one() : Fq2::zero();
    const Fq2 r = final_exponentiation(L, m);
    const bool eq = all_lanes(L, r == reinterpret_cast<const Fq2 *>(alpha_beta)[slot_index(L.t)]);
    if (in_range && L.live && L.t == 0) verdict[i] = bad ? bad : ((eq && !zero) ? 1 : 0);
}

void zk_launch_miller_lanes(cudaStream_t st, size_t n, const void *a, const void *acc, const void *c, const void *coef_b, const void *gamma, int gamma_inf,
                            const void *delta, int delta_inf, const uint8_t *status, void *f) {
    const size_t warps = (n + GROUPS_PER_WARP - 1) / GROUPS_PER_WARP;
    const unsigned g = (unsigned)((warps * 32 + 127) / 128);
    k_miller_lanes<<<g, 128, 0, st>>>(n, (const G1A *)a, (const G1A *)acc, (const G1A *)c, (const zkpair::LineCoeff *)coef_b, (const zkpair::LineCoeff *)gamma, gamma_inf,
                                             (const zkpair::LineCoeff *)delta, delta_inf, status, (zkpair::Fq12 *)f);
}
void zk_launch_verify_final_lanes(cudaStream_t st, size_t n, const void *f, const void *alpha_beta, const uint8_t *status, uint8_t *verdict) {
    const size_t warps = (n + GROUPS_PER_WARP - 1) / GROUPS_PER_WARP;
    const unsigned g = (unsigned)((warps * 32 + 127) / 128);
    k_verify_final_lanes<<<g, 128, 0, st>>>(n, (const zkpair::Fq12 *)f, (const zkpair::Fq12 *)alpha_beta, status, verdict);
}
