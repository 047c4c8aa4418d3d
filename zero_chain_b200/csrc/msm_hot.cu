// Hot translation unit: G1 MSM driver + kernels with Montgomery products / point additions fully
// inlined (ZK_HOT).  Everything else in the library calls them as functions (field.cuh, inlining policy).
#define ZK_HOT 1
#include "msm_driver.cuh"
#ifdef ZK_EXPERIMENTS
#include "field_wide.cuh"
#endif

using namespace zkmsm;

int zk_msm_run_g1(zk_ctx *ctx, const zk_bases *b, const uint32_t *d_scalars, size_t n, size_t batch) { return msm_run_t<Fq>(ctx, b, d_scalars, n, batch); }
int zk_build_tables_g1(zk_ctx *ctx, zk_bases *b) { return build_tables_t<Fq>(ctx, b); }
int zk_encode_results_g1(zk_ctx *ctx, size_t count, int compressed, uint8_t *d_out) { return encode_results_t<Fq>(ctx, count, compressed, d_out); }

// modmul roofline calibration: 4 independent chains per thread, register resident
template <class T>
__global__ void k_bench_modmul(int iters, T *sink) {
    T a[4], b = T::one();
    for (int k = 0; k < 4; k++) { a[k] = T::one(); a[k].l[0] += threadIdx.x + k + 1; }
    b.l[1] ^= blockIdx.x + 7;
    a[0] = T::reduce_once(a[0]); b = T::reduce_once(b);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) a[k] = a[k] * b;
    }
    T r = a[0] + a[1] + a[2] + a[3];
    if (r.l[0] == 0x12345678u && r.l[1] == 0x9abcdef0u) sink[0] = r;   // keep the work alive
}
#ifdef ZK_EXPERIMENTS
// experiment: separated multiply / reduce (field_wide.cuh) against the interleaved product, 4 chains per thread
template <int MODE>
__global__ void k_bench_wide(int iters, Fq *sink) {
    Fq a[4], b = Fq::one();
    for (int k = 0; k < 4; k++) { a[k] = Fq::one(); a[k].l[0] += threadIdx.x + k + 1; }
    b.l[1] ^= blockIdx.x + 7;
    a[0] = Fq::reduce_once(a[0]); b = Fq::reduce_once(b);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (MODE == 0) a[k] = a[k] * b;
            else if (MODE == 1) a[k] = zkwide::mul_sep(a[k], b);
            else if (MODE == 2) a[k] = a[k] * a[k];
            else if (MODE == 3) a[k] = zkwide::sqr_sep(a[k]);
            else if (MODE == 4) a[k] = a[k] * b - b * a[(k + 1) & 3];
            else a[k] = zkwide::mul_sub_mul(a[k], b, b, a[(k + 1) & 3]);
        }
    }
    Fq r = a[0] + a[1] + a[2] + a[3];
    if (r.l[0] == 0x12345678u && r.l[1] == 0x9abcdef0u) sink[0] = r;
}
#endif
void zk_launch_bench_modmul(int field, int blocks, int threads, int iters, void *sink, cudaStream_t st) {
    if (field == 0) k_bench_modmul<Fq><<<blocks, threads, 0, st>>>(iters, (Fq *)sink);
    else if (field == 1) k_bench_modmul<Fr><<<blocks, threads, 0, st>>>(iters, (Fr *)sink);
#ifdef ZK_EXPERIMENTS
    else if (field == 10) k_bench_wide<0><<<blocks, threads, 0, st>>>(iters, (Fq *)sink);
    else if (field == 11) k_bench_wide<1><<<blocks, threads, 0, st>>>(iters, (Fq *)sink);
    else if (field == 12) k_bench_wide<2><<<blocks, threads, 0, st>>>(iters, (Fq *)sink);
    else if (field == 13) k_bench_wide<3><<<blocks, threads, 0, st>>>(iters, (Fq *)sink);
    else if (field == 14) k_bench_wide<4><<<blocks, threads, 0, st>>>(iters, (Fq *)sink);
    else k_bench_wide<5><<<blocks, threads, 0, st>>>(iters, (Fq *)sink);
#endif
}
