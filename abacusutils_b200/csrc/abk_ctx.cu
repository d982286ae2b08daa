// Context, error reporting and small device utilities of libabk.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "abk_common.cuh"

static thread_local char g_err[512] = "";

void abk_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *abk_last_error(void) { return g_err; }
extern "C" int abk_version(void) { return ABK_VERSION; }

extern "C" int abk_ctx_create(int device, abk_ctx **out)
{
    ABK_REQUIRE(out != nullptr, "abk_ctx_create: null output pointer");
    int ndev = 0;
    ABK_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    ABK_REQUIRE(device >= 0 && device < ndev, "abk_ctx_create: device %d out of range (%d visible)", device, ndev);
    abk_device_guard guard(device);  // the caller's current device is restored on return
    cudaDeviceProp prop;
    ABK_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        abk_set_error("abk_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                      device, prop.major, prop.minor);
        return ABK_ERR_INVALID;
    }
    abk_ctx *c = new abk_ctx();
    c->device = device;
    c->stream = nullptr;
    c->num_sms = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    c->launches = 0;
    c->tile_capacity = 0;
    c->scheme = 0;
    c->bin_no_sym = 0;
    c->wscale = 1.0f;
    c->d_scalars = nullptr;
    c->prof_on = 0;
    c->prof_recs = nullptr;
    c->prof_n = c->prof_cap = 0;
    c->prof_pool = nullptr;
    c->prof_pool_n = c->prof_pool_cap = 0;
    ABK_CHECK_CUDA(cudaMalloc(&c->d_scalars, 64 * sizeof(unsigned long long)));
    ABK_CHECK_CUDA(cudaMemset(c->d_scalars, 0, 64 * sizeof(unsigned long long)));
    *out = c;
    return ABK_OK;
}

extern "C" int abk_ctx_destroy(abk_ctx *ctx)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    if (!ctx) return ABK_OK;
    abk_device_guard guard(ctx->device);
    if (ctx->d_scalars) cudaFree(ctx->d_scalars);
    delete ctx;
    return ABK_OK;
}

// ---- per-kernel timing with CUDA events on the launch stream ------------------------------------
static const char *const k_names[ABK_K_COUNT] = {
    "wrap_inplace", "partition_hist", "partition_scatter", "scan", "tsc_bucket_hist", "tsc_bucket_scatter",
    "tsc_tile_deposit", "tsc_naive_deposit", "normalize_field", "cufft", "field_fft_finish", "raw_power",
    "power_bin", "add_planes", "transpose_pack", "misc"};

static cudaEvent_t prof_get_event(abk_ctx *ctx)
{
    if (ctx->prof_pool_n > 0) return ctx->prof_pool[--ctx->prof_pool_n];
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void abk_prof_begin(abk_ctx *ctx, int id)
{
    if (ctx->prof_n == ctx->prof_cap) {
        const int cap = ctx->prof_cap ? 2 * ctx->prof_cap : 256;
        ctx->prof_recs = (abk_prof_rec *)realloc(ctx->prof_recs, sizeof(abk_prof_rec) * cap);
        ctx->prof_cap = cap;
    }
    abk_prof_rec &r = ctx->prof_recs[ctx->prof_n++];
    r.id = id;
    r.a = prof_get_event(ctx);
    r.b = prof_get_event(ctx);
    cudaEventRecord(r.a, ctx->stream);
}

void abk_prof_end(abk_ctx *ctx)
{
    cudaEventRecord(ctx->prof_recs[ctx->prof_n - 1].b, ctx->stream);
}

extern "C" int abk_ctx_profile_enable(abk_ctx *ctx, int on)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx != nullptr, "null context");
    ctx->prof_on = on ? 1 : 0;
    return ABK_OK;
}

extern "C" int abk_kernel_count(void) { return ABK_K_COUNT; }
extern "C" const char *abk_kernel_name(int id) { return (id >= 0 && id < ABK_K_COUNT) ? k_names[id] : ""; }

// Synchronises the stream, adds the elapsed time of every recorded launch to ms_h[id] and its
// count to n_h[id] (arrays of abk_kernel_count() entries, NOT zeroed here), and clears the records.
extern "C" int abk_ctx_profile_collect(abk_ctx *ctx, double *ms_h, int64_t *n_h)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && ms_h && n_h, "abk_ctx_profile_collect: null argument");
    ABK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ctx->prof_n; i++) {
        abk_prof_rec &r = ctx->prof_recs[i];
        float ms = 0.0f;
        ABK_CHECK_CUDA(cudaEventSynchronize(r.b));
        ABK_CHECK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        ms_h[r.id] += ms;
        n_h[r.id] += 1;
        if (ctx->prof_pool_n + 2 > ctx->prof_pool_cap) {
            const int cap = ctx->prof_pool_cap ? 2 * ctx->prof_pool_cap : 512;
            ctx->prof_pool = (cudaEvent_t *)realloc(ctx->prof_pool, sizeof(cudaEvent_t) * cap);
            ctx->prof_pool_cap = cap;
        }
        ctx->prof_pool[ctx->prof_pool_n++] = r.a;
        ctx->prof_pool[ctx->prof_pool_n++] = r.b;
    }
    ctx->prof_n = 0;
    return ABK_OK;
}

extern "C" int abk_ctx_set_stream(abk_ctx *ctx, void *stream)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx != nullptr, "null context");
    ctx->stream = (cudaStream_t)stream;
    return ABK_OK;
}

extern "C" int abk_ctx_sync(abk_ctx *ctx)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx != nullptr, "null context");
    ABK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return ABK_OK;
}

extern "C" int64_t abk_ctx_launch_count(abk_ctx *ctx) { return ctx ? ctx->launches : -1; }

extern "C" int abk_ctx_set_scheme(abk_ctx *ctx, int scheme)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx != nullptr, "null context");
    ABK_REQUIRE(scheme == 0 || scheme == 1, "unknown mass-assignment scheme %d (0 = TSC, 1 = CIC)", scheme);
    ctx->scheme = scheme;
    return ABK_OK;
}

extern "C" int abk_ctx_set_weight_scale(abk_ctx *ctx, double scale)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx != nullptr, "null context");
    ABK_REQUIRE(scale == scale && scale != 0.0, "weight scale must be a non-zero number");
    ctx->wscale = (float)scale;
    return ABK_OK;
}

extern "C" int abk_ctx_set_tile_capacity(abk_ctx *ctx, int capacity)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx != nullptr, "null context");
    const int cap = capacity & 0xffff;
    ABK_REQUIRE(cap == 0 || (cap >= 256 && cap <= 12288), "tile capacity %d out of range", cap);
    ctx->tile_capacity = cap;
    ctx->bin_no_sym = (capacity >> 19) & 1;  // bit 19: disable the mirror-symmetric binning kernel (experiments)
    return ABK_OK;
}

// ------------------------------------------------------------------------------------------
// Inclusive scan of uint32 (tile histograms -> bucket boundaries).  Three-level recursive
// block scan; the arrays are small (<= a few 10^7 entries) next to the particle data.
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_kernel(uint32_t *data, int64_t n, uint32_t *block_sums)
{
    __shared__ uint32_t warp_tot[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t run = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) {
        const int64_t i = base + q;
        run += (i < n) ? data[i] : 0u;
        v[q] = run;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = warp_tot[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += t;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    const uint32_t prefix = (inc - run) + (wid ? warp_tot[wid - 1] : 0u);
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) {
        const int64_t i = base + q;
        if (i < n) data[i] = v[q] + prefix;
    }
    if (block_sums && threadIdx.x == SCAN_THREADS - 1) block_sums[blockIdx.x] = prefix + run;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(uint32_t *data, int64_t n, const uint32_t *block_incl)
{
    if (blockIdx.x == 0) return;
    const uint32_t add = block_incl[blockIdx.x - 1];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) {
        const int64_t i = base + q;
        if (i < n) data[i] += add;
    }
}

size_t abk_scan_tmp_bytes(int64_t n)
{
    size_t total = 0;
    int64_t m = n;
    while (m > SCAN_BLOCK) {
        m = (m + SCAN_BLOCK - 1) / SCAN_BLOCK;
        total += abk_align_up((size_t)m * sizeof(uint32_t), 256);
    }
    return total + 256;
}

int abk_inclusive_scan_u32(abk_ctx *ctx, uint32_t *data, int64_t n, void *tmp)
{
    if (n <= 0) return ABK_OK;
    const int64_t nblocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    uint32_t *sums = (nblocks > 1) ? (uint32_t *)tmp : nullptr;
    ABK_LAUNCH(ctx, ABK_K_SCAN, scan_block_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, ctx->stream>>>(data, n, sums));
    if (nblocks > 1) {
        char *next = (char *)tmp + abk_align_up((size_t)nblocks * sizeof(uint32_t), 256);
        int rc = abk_inclusive_scan_u32(ctx, sums, nblocks, next);
        if (rc) return rc;
        ABK_LAUNCH(ctx, ABK_K_SCAN, scan_add_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, ctx->stream>>>(data, n, sums));
    }
    return ABK_OK;
}
