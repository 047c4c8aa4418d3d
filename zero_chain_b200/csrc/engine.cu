// libzkb200: execution context, MSM driver and diagnostics (host side of the C ABI, include/zkb200.h).
// The reference's counterpart is bellman's `multiexp` + `Worker` CPU pool (un-vendored);
// here the "pool" is one CUDA stream per context and the schedule documented in msm.cuh.
#include "internal.h"
#include "msm.cuh"
#include "codec.cuh"
#include <type_traits>

using namespace zkmsm;

static thread_local char g_err[512] = "";
void zk_set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char *zk_last_error(void) { return g_err; }
extern "C" const char *zk_version(void) { return "zkb200 0.1 (sm_100a)"; }
extern "C" int zk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int zk_use_device(zk_ctx *ctx) { ZK_CUDA(cudaSetDevice(ctx->device)); return ZK_OK; }

extern "C" int zk_ctx_create(int device, void *stream, zk_ctx **out) {
    if (!out) { zk_set_error("zk_ctx_create: out is NULL"); return ZK_ERR_INVALID; }
    int n = zk_device_count();
    if (n == 0) { zk_set_error("no CUDA device: libzkb200 has no CPU fallback"); return ZK_ERR_CUDA; }
    if (device < 0 || device >= n) { zk_set_error("device %d out of range (%d devices)", device, n); return ZK_ERR_INVALID; }
    ZK_CUDA(cudaSetDevice(device));
    zk_ctx *c = new zk_ctx();
    c->device = device;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { ZK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    cudaDeviceProp prop;
    ZK_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    ZK_CUDA(cudaMalloc(&c->d_err, 16 * sizeof(int)));
    ZK_CUDA(cudaMemsetAsync(c->d_err, 0, 16 * sizeof(int), c->stream));
    c->h_pinned_cap = 1 << 20;
    ZK_CUDA(cudaMallocHost(&c->h_pinned, c->h_pinned_cap));
    *out = c;
    return ZK_OK;
}
extern "C" void zk_ctx_destroy(zk_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = {&c->scalars, &c->digits, &c->tile_hist, &c->tile_off, &c->sizes, &c->bucket_off, &c->task_off, &c->scan_scratch,
                      &c->sorted, &c->partials, &c->buckets, &c->red_part, &c->red_x, &c->result, &c->out_bytes, &c->stage_a, &c->stage_b,
                      &c->stage_c, &c->ntt_tw, &c->ntt_tmp, &c->g_a, &c->g_b, &c->g_c, &c->g_h, &c->g_scal, &c->g_misc};
    for (DevBuf *b : bufs) b->release();
    if (c->d_err) cudaFree(c->d_err);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}
extern "C" int zk_ctx_sync(zk_ctx *c) { ZK_TRY(zk_use_device(c)); ZK_CUDA(cudaStreamSynchronize(c->stream)); return ZK_OK; }
extern "C" void *zk_ctx_stream(zk_ctx *c) { return (void *)c->stream; }

static int check_err_flag(zk_ctx *ctx) {
    int e[2] = {0, 0};
    ZK_CUDA(cudaMemcpyAsync(e, ctx->d_err, sizeof(e), cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (e[0] || e[1]) {
        ZK_CUDA(cudaMemsetAsync(ctx->d_err, 0, 2 * sizeof(int), ctx->stream));
        if (e[0]) { zk_set_error("scalar not canonical (>= r)"); return ZK_ERR_NOT_CANONICAL; }
        zk_set_error("point decoding failed (GroupDecodingError %d)", e[1]);
        return e[1] == zkcodec::DEC_INFINITY ? ZK_ERR_UNEXPECTED_IDENTITY : ZK_ERR_DECODE;
    }
    return ZK_OK;
}

// ---- bases ------------------------------------------------------------------------------------------
static int pick_window(size_t n) {
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    int c = lg - 2;
    if (c < 5) c = 5;
    if (c > 16) c = 16;
    return c;
}
__global__ void k_any_inf(const uint32_t *limbs, size_t n, int words, int *err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o = 0;
    for (int k = 0; k < words; k++) o |= limbs[i * words + k];
    if (!o) atomicCAS(err + 1, 0, zkcodec::DEC_INFINITY);
}
template <class F>
static int build_tables(zk_ctx *ctx, zk_bases *b) {
    unsigned thr = 128, blk = (unsigned)((b->n + thr * PRE_K - 1) / (thr * PRE_K));
    k_precompute<F><<<blk, thr, 0, ctx->stream>>>((Affine<F> *)b->d_tbl, (uint32_t)b->n, b->c, b->W);
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}
// device-resident variant used by the Groth16 CRS loader: d_points already holds n affine points
int zk_bases_from_device(zk_ctx *ctx, int group, const void *d_points, size_t n, int window_bits, int precompute, zk_bases **out) {
    if (group != 1 && group != 2) { zk_set_error("group must be 1 or 2"); return ZK_ERR_INVALID; }
    if (n == 0 || n >= ((size_t)1 << 27)) { zk_set_error("unsupported base count %zu", n); return ZK_ERR_INVALID; }
    ZK_TRY(zk_use_device(ctx));
    zk_bases *b = new zk_bases();
    b->group = group; b->device = ctx->device; b->n = n;
    b->c = window_bits > 0 ? window_bits : pick_window(n);
    if (b->c < 2 || b->c > 16) { delete b; zk_set_error("window_bits must be in [2,16]"); return ZK_ERR_INVALID; }
    b->W = 255 / b->c + 1;
    b->tables = precompute != 0;
    size_t psz = group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    size_t rows = b->tables ? b->W : 1;
    cudaError_t e = cudaMalloc(&b->d_tbl, rows * n * psz);
    if (e != cudaSuccess) { delete b; zk_set_error("cudaMalloc tables (%zu B): %s", rows * n * psz, cudaGetErrorString(e)); return ZK_ERR_CUDA; }
    ZK_CUDA(cudaMemcpyAsync(b->d_tbl, d_points, n * psz, cudaMemcpyDeviceToDevice, ctx->stream));
    k_any_inf<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t *)b->d_tbl, n, (int)(psz / 4), ctx->d_err);
    int r = check_err_flag(ctx);
    if (r) { zk_bases_free(b); return r; }
    if (b->tables) {
        r = group == 1 ? build_tables<Fq>(ctx, b) : build_tables<Fq2>(ctx, b);
        if (r) { zk_bases_free(b); return r; }
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = b;
    return ZK_OK;
}
extern "C" int zk_bases_upload(zk_ctx *ctx, int group, const uint64_t *limbs, size_t n, int window_bits, int precompute, zk_bases **out) {
    if (!ctx || !limbs || !out) { zk_set_error("zk_bases_upload: NULL argument"); return ZK_ERR_INVALID; }
    if (group != 1 && group != 2) { zk_set_error("group must be 1 or 2"); return ZK_ERR_INVALID; }
    ZK_TRY(zk_use_device(ctx));
    size_t psz = group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    ZK_TRY(ctx->stage_a.reserve(n * psz));
    ZK_CUDA(cudaMemcpyAsync(ctx->stage_a.p, limbs, n * psz, cudaMemcpyHostToDevice, ctx->stream));
    return zk_bases_from_device(ctx, group, ctx->stage_a.p, n, window_bits, precompute, out);
}
extern "C" void zk_bases_free(zk_bases *b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->d_tbl) cudaFree(b->d_tbl);
    delete b;
}
extern "C" size_t zk_bases_len(const zk_bases *b) { return b ? b->n : 0; }
extern "C" int zk_bases_window_bits(const zk_bases *b) { return b ? b->c : 0; }

// ---- MSM driver ------------------------------------------------------------------------------------------
template <class F>
static int msm_run_t(zk_ctx *ctx, const zk_bases *b, const uint32_t *d_scalars, size_t n, size_t batch) {
    cudaStream_t st = ctx->stream;
    const int c = b->c, W = b->W, nbins = 1 << (c - 1);
    // sort domains: with tables one domain per batch item holding all W windows; without tables one per window
    const bool tables = b->tables;
    if (!tables && batch != 1) { zk_set_error("batched MSM needs precomputed tables"); return ZK_ERR_INVALID; }
    const size_t n_dom = tables ? batch : (size_t)W;
    const uint64_t e_dom = tables ? (uint64_t)n * W : (uint64_t)n;
    const size_t E = (size_t)n * W * batch;
    if (E >= ((size_t)1 << 31)) { zk_set_error("MSM too large for 31-bit entry payloads (n*W*batch = %zu)", E); return ZK_ERR_INVALID; }
    const int tiles = (int)((e_dom + TILE - 1) / TILE);
    const size_t NB = n_dom * nbins;
    const size_t t_max = E / TASK_LEN + NB + 1;
    const size_t pt = sizeof(XYZZ<F>);
    ZK_TRY(ctx->digits.reserve(E * 4));
    ZK_TRY(ctx->tile_hist.reserve(n_dom * tiles * (size_t)nbins * 2));
    ZK_TRY(ctx->tile_off.reserve(n_dom * tiles * (size_t)nbins * 4));
    ZK_TRY(ctx->sizes.reserve((NB + 1) * 4));
    ZK_TRY(ctx->bucket_off.reserve((NB + 1) * 4));
    ZK_TRY(ctx->task_off.reserve((NB + 1) * 4));
    ZK_TRY(ctx->scan_scratch.reserve((2 * (NB / SCAN_B + 8) + 4096) * 4));
    ZK_TRY(ctx->sorted.reserve(E * 4));
    ZK_TRY(ctx->partials.reserve(t_max * pt));
    ZK_TRY(ctx->buckets.reserve(NB * pt));
    const int n_bits = c;                      // digit values d in [1, 2^(c-1)] need c bits
    const int n_slices = (nbins + RED_SLICE - 1) / RED_SLICE;
    ZK_TRY(ctx->red_part.reserve(n_dom * n_bits * (size_t)n_slices * pt));
    ZK_TRY(ctx->red_x.reserve(n_dom * n_bits * pt));
    ZK_TRY(ctx->result.reserve((n_dom + batch + 1) * pt));

    uint32_t *digits = ctx->digits.as<uint32_t>();
    {   // 1. digits: grid.y = batch item, layout [batch][W][n]
        dim3 g((unsigned)((n + 255) / 256), (unsigned)batch);
        k_msm_digits<<<g, 256, 0, st>>>(d_scalars, (uint32_t)n, c, W, digits, ctx->d_err);
    }
    // 2. counting sort per domain
    size_t smem = (size_t)nbins * 4;
    if (smem > 48 * 1024) {
        ZK_CUDA(cudaFuncSetAttribute(k_tile_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ZK_CUDA(cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 gs((unsigned)tiles, (unsigned)n_dom);
    k_tile_hist<<<gs, SORT_THREADS, smem, st>>>(digits, e_dom, nbins, ctx->tile_hist.as<uint16_t>(), tiles);
    k_col_scan<<<(unsigned)((NB + 255) / 256), 256, 0, st>>>(ctx->tile_hist.as<uint16_t>(), ctx->tile_off.as<uint32_t>(), ctx->sizes.as<uint32_t>(),
                                                            nbins, tiles, (int)n_dom);
    exclusive_scan<false>(ctx->sizes.as<uint32_t>(), ctx->bucket_off.as<uint32_t>(), NB, ctx->scan_scratch.as<uint32_t>(), st);
    exclusive_scan<true>(ctx->sizes.as<uint32_t>(), ctx->task_off.as<uint32_t>(), NB, ctx->scan_scratch.as<uint32_t>(), st);
    k_scatter<<<gs, SORT_THREADS, smem, st>>>(digits, e_dom, nbins, ctx->tile_off.as<uint32_t>(), ctx->bucket_off.as<uint32_t>(),
                                              ctx->sorted.as<uint32_t>(), tiles);
    // 3. accumulate + combine.  The payload of an entry is its position in the domain = [w][i] index;
    //    with tables that is the table index when n == b->n (checked by the callers).
    XYZZ<F> *partials = ctx->partials.as<XYZZ<F>>(), *buckets = ctx->buckets.as<XYZZ<F>>();
    if constexpr (std::is_same<F, Fq>::value)
        zk_launch_accumulate_g1(b->d_tbl, ctx->sorted.as<uint32_t>(), ctx->bucket_off.as<uint32_t>(), ctx->task_off.as<uint32_t>(), (uint32_t)NB,
                                partials, t_max, st);
    else
        k_accumulate<F><<<(unsigned)((t_max + 127) / 128), 128, 0, st>>>((const Affine<F> *)b->d_tbl, ctx->sorted.as<uint32_t>(),
                                                                         ctx->bucket_off.as<uint32_t>(), ctx->task_off.as<uint32_t>(), (uint32_t)NB, partials);
    size_t sm_comb = 4 * 32 * pt;
    if (sm_comb > 48 * 1024) ZK_CUDA(cudaFuncSetAttribute(k_combine<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_comb));
    k_combine<F><<<(unsigned)((NB * 32 + 127) / 128), 128, sm_comb, st>>>(partials, ctx->task_off.as<uint32_t>(), (uint32_t)NB, buckets);
    // 4. bucket reduction per domain
    size_t sm_red = RED_T * pt;
    if (sm_red > 48 * 1024) {
        ZK_CUDA(cudaFuncSetAttribute(k_bit_sums<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_red));
        ZK_CUDA(cudaFuncSetAttribute(k_sum_points<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_red));
    }
    XYZZ<F> *part = ctx->red_part.as<XYZZ<F>>(), *X = ctx->red_x.as<XYZZ<F>>(), *R = ctx->result.as<XYZZ<F>>();
    k_bit_sums<F><<<dim3((unsigned)n_slices, (unsigned)n_bits, (unsigned)n_dom), RED_T, sm_red, st>>>(buckets, nbins, n_slices, n_bits, part);
    k_sum_points<F><<<(unsigned)(n_dom * n_bits), RED_T, sm_red, st>>>(part, n_slices, X);
    if (tables) {
        k_finish_bits<F><<<(unsigned)((n_dom + 63) / 64), 64, 0, st>>>(X, n_bits, (int)n_dom, R);
    } else {
        k_finish_bits<F><<<(unsigned)((n_dom + 63) / 64), 64, 0, st>>>(X, n_bits, (int)n_dom, R + 1);
        k_horner_windows<F><<<1, 32, 0, st>>>(R + 1, W, c, R);
    }
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}
int zk_msm_run(zk_ctx *ctx, const zk_bases *b, const void *d_scalars, size_t n, size_t batch) {
    if (!ctx || !b || !d_scalars) { zk_set_error("zk_msm: NULL argument"); return ZK_ERR_INVALID; }
    if (b->device != ctx->device) { zk_set_error("bases live on device %d, context on %d", b->device, ctx->device); return ZK_ERR_INVALID; }
    if (n == 0 || batch == 0) { zk_set_error("empty MSM"); return ZK_ERR_INVALID; }
    if (b->tables ? n != b->n : n > b->n) {
        zk_set_error("scalar count %zu does not match the %zu bases (SynthesisError::AssignmentMissing)", n, b->n);
        return ZK_ERR_ASSIGNMENT_MISSING;
    }
    ZK_TRY(zk_use_device(ctx));
    return b->group == 1 ? msm_run_t<Fq>(ctx, b, (const uint32_t *)d_scalars, n, batch) : msm_run_t<Fq2>(ctx, b, (const uint32_t *)d_scalars, n, batch);
}
int zk_encode_results(zk_ctx *ctx, int group, size_t count, int compressed, uint8_t *out_host) {
    size_t per = (group == 1 ? 96 : 192) / (compressed ? 2 : 1);
    ZK_TRY(ctx->out_bytes.reserve(count * per));
    if (count * per > ctx->h_pinned_cap) { zk_set_error("result batch too large"); return ZK_ERR_INVALID; }
    unsigned blk = (unsigned)((count + 63) / 64);
    if (group == 1) zkcodec::k_encode_xyzz<Fq><<<blk, 64, 0, ctx->stream>>>(ctx->result.as<G1XYZZ>(), (int)count, compressed, ctx->out_bytes.as<uint8_t>());
    else zkcodec::k_encode_xyzz<Fq2><<<blk, 64, 0, ctx->stream>>>(ctx->result.as<G2XYZZ>(), (int)count, compressed, ctx->out_bytes.as<uint8_t>());
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->out_bytes.p, count * per, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_TRY(check_err_flag(ctx));   // synchronises the stream
    memcpy(out_host, ctx->h_pinned, count * per);
    return ZK_OK;
}
extern "C" int zk_msm_batch_device(zk_ctx *ctx, const zk_bases *b, const void *d_scalars, size_t n, size_t batch, uint8_t *out) {
    if (!out) { zk_set_error("zk_msm: out is NULL"); return ZK_ERR_INVALID; }
    ZK_TRY(zk_msm_run(ctx, b, d_scalars, n, batch));
    return zk_encode_results(ctx, b->group, batch, 0, out);
}
extern "C" int zk_msm_device(zk_ctx *ctx, const zk_bases *b, const void *d_scalars, size_t n, uint8_t *out) {
    return zk_msm_batch_device(ctx, b, d_scalars, n, 1, out);
}
extern "C" int zk_msm(zk_ctx *ctx, const zk_bases *b, const uint64_t *scalars, size_t n, uint8_t *out) {
    if (!ctx || !scalars) { zk_set_error("zk_msm: NULL argument"); return ZK_ERR_INVALID; }
    ZK_TRY(zk_use_device(ctx));
    ZK_TRY(ctx->scalars.reserve(n * 32));
    ZK_CUDA(cudaMemcpyAsync(ctx->scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return zk_msm_device(ctx, b, ctx->scalars.p, n, out);
}
extern "C" size_t zk_partial_size(int group) { return group == 1 ? sizeof(G1XYZZ) : sizeof(G2XYZZ); }
extern "C" int zk_msm_partial_device(zk_ctx *ctx, const zk_bases *b, const void *d_scalars, size_t n, void *d_partial_out) {
    if (!d_partial_out) { zk_set_error("zk_msm_partial_device: out is NULL"); return ZK_ERR_INVALID; }
    ZK_TRY(zk_msm_run(ctx, b, d_scalars, n, 1));
    ZK_CUDA(cudaMemcpyAsync(d_partial_out, ctx->result.p, zk_partial_size(b->group), cudaMemcpyDeviceToDevice, ctx->stream));
    return check_err_flag(ctx);
}
template <class F>
__global__ void k_fold_serial(const XYZZ<F> *in, int n, XYZZ<F> *out) {
    if (threadIdx.x | blockIdx.x) return;
    XYZZ<F> r = in[0];
    for (int i = 1; i < n; i++) r.add(in[i]);
    out[0] = r;
}
extern "C" int zk_points_fold(zk_ctx *ctx, int group, const void *d_partials, size_t count, uint8_t *out) {
    if (!ctx || !d_partials || !out || count == 0) { zk_set_error("zk_points_fold: bad argument"); return ZK_ERR_INVALID; }
    ZK_TRY(zk_use_device(ctx));
    ZK_TRY(ctx->result.reserve(4 * sizeof(G2XYZZ)));
    if (group == 1) k_fold_serial<Fq><<<1, 32, 0, ctx->stream>>>((const G1XYZZ *)d_partials, (int)count, ctx->result.as<G1XYZZ>());
    else k_fold_serial<Fq2><<<1, 32, 0, ctx->stream>>>((const G2XYZZ *)d_partials, (int)count, ctx->result.as<G2XYZZ>());
    ZK_CUDA(cudaGetLastError());
    return zk_encode_results(ctx, group, 1, 0, out);
}

// ---- utilities -----------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_scalar_mul_many(const Affine<F> *base, const uint32_t *scalars, size_t n, Affine<F> *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
    for (int j = 0; j < 8; j++) k[j] = scalars[i * 8 + j];
    out[i] = scalar_mul(XYZZ<F>::from_affine(base[0]), k).to_affine();
}
extern "C" int zk_scalar_mul_many(zk_ctx *ctx, int group, const uint64_t *base, const uint64_t *scalars, size_t n, uint64_t *out) {
    if (!ctx || !base || !scalars || !out) { zk_set_error("zk_scalar_mul_many: NULL argument"); return ZK_ERR_INVALID; }
    if (group != 1 && group != 2) { zk_set_error("group must be 1 or 2"); return ZK_ERR_INVALID; }
    ZK_TRY(zk_use_device(ctx));
    size_t psz = group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
    ZK_TRY(ctx->stage_a.reserve(psz)); ZK_TRY(ctx->stage_b.reserve(n * 32)); ZK_TRY(ctx->stage_c.reserve(n * psz));
    ZK_CUDA(cudaMemcpyAsync(ctx->stage_a.p, base, psz, cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMemcpyAsync(ctx->stage_b.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    unsigned blk = (unsigned)((n + 127) / 128);
    if (group == 1) k_scalar_mul_many<Fq><<<blk, 128, 0, ctx->stream>>>(ctx->stage_a.as<G1Affine>(), ctx->stage_b.as<uint32_t>(), n, ctx->stage_c.as<G1Affine>());
    else k_scalar_mul_many<Fq2><<<blk, 128, 0, ctx->stream>>>(ctx->stage_a.as<G2Affine>(), ctx->stage_b.as<uint32_t>(), n, ctx->stage_c.as<G2Affine>());
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaMemcpyAsync(out, ctx->stage_c.p, n * psz, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

template <class T>
__global__ void k_field_op(int op, const T *a, const T *b, size_t n, T *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T x = a[i], y = b ? b[i] : x, r;
    switch (op) {
    case 0: r = x * y; break;
    case 1: r = x + y; break;
    case 2: r = x - y; break;
    case 3: r = x.sqr(); break;
    case 4: r = x.inverse(); break;
    case 5: r = T::canonical_lt_mod(x) ? T::from_canonical(x) : T::zero(); break;
    default: r = x.to_canonical(); break;
    }
    out[i] = r;
}
extern "C" int zk_field_op(zk_ctx *ctx, int field, int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out) {
    if (!ctx || !a || !out || op < 0 || op > 6 || (field != 0 && field != 1)) { zk_set_error("zk_field_op: bad argument"); return ZK_ERR_INVALID; }
    ZK_TRY(zk_use_device(ctx));
    size_t sz = field == 0 ? 48 : 32;
    ZK_TRY(ctx->stage_a.reserve(n * sz)); ZK_TRY(ctx->stage_b.reserve(n * sz)); ZK_TRY(ctx->stage_c.reserve(n * sz));
    ZK_CUDA(cudaMemcpyAsync(ctx->stage_a.p, a, n * sz, cudaMemcpyHostToDevice, ctx->stream));
    if (b) ZK_CUDA(cudaMemcpyAsync(ctx->stage_b.p, b, n * sz, cudaMemcpyHostToDevice, ctx->stream));
    unsigned blk = (unsigned)((n + 127) / 128);
    if (field == 0) k_field_op<Fq><<<blk, 128, 0, ctx->stream>>>(op, ctx->stage_a.as<Fq>(), b ? ctx->stage_b.as<Fq>() : nullptr, n, ctx->stage_c.as<Fq>());
    else k_field_op<Fr><<<blk, 128, 0, ctx->stream>>>(op, ctx->stage_a.as<Fr>(), b ? ctx->stage_b.as<Fr>() : nullptr, n, ctx->stage_c.as<Fr>());
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaMemcpyAsync(out, ctx->stage_c.p, n * sz, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

extern "C" int zk_bench_modmul(zk_ctx *ctx, int field, int blocks, int threads, int iters, double *per_s, double *ms_out) {
    if (!ctx || !per_s) { zk_set_error("zk_bench_modmul: NULL argument"); return ZK_ERR_INVALID; }
    ZK_TRY(zk_use_device(ctx));
    ZK_TRY(ctx->stage_a.reserve(64));
    cudaEvent_t e0, e1;
    ZK_CUDA(cudaEventCreate(&e0)); ZK_CUDA(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {
        ZK_CUDA(cudaEventRecord(e0, ctx->stream));
        zk_launch_bench_modmul(field, blocks, threads, iters, ctx->stage_a.p, ctx->stream);
        ZK_CUDA(cudaEventRecord(e1, ctx->stream));
        ZK_CUDA(cudaEventSynchronize(e1));
    }
    float ms = 0;
    ZK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *per_s = (double)blocks * threads * iters * 4.0 / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return ZK_OK;
}
