// Hot translation unit: kernels whose Montgomery products / point additions are fully inlined
// (ZK_HOT).  Everything else in the library calls them as functions (see field.cuh, inlining policy).
#define ZK_HOT 1
#include "internal.h"
#include "msm_accum.cuh"

using namespace zkmsm;

void zk_launch_accumulate_g1(const void *bases, const uint32_t *sorted, const uint32_t *bucket_off, const uint32_t *task_off,
                             uint32_t n_buckets, void *partials, size_t t_max, cudaStream_t st) {
    k_accumulate<Fq><<<(unsigned)((t_max + 127) / 128), 128, 0, st>>>((const G1Affine *)bases, sorted, bucket_off, task_off, n_buckets,
                                                                     (G1XYZZ *)partials);
}

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
void zk_launch_bench_modmul(int field, int blocks, int threads, int iters, void *sink, cudaStream_t st) {
    if (field == 0) k_bench_modmul<Fq><<<blocks, threads, 0, st>>>(iters, (Fq *)sink);
    else k_bench_modmul<Fr><<<blocks, threads, 0, st>>>(iters, (Fr *)sink);
}
