// FFT stage: power_spectrum.py:980,986,1059 call scipy.fft.rfftn (pocketfft); here the same
// unnormalised forward R2C transform runs in place through cuFFT (BASELINE.json north_star:
// "the real-to-complex 3D FFT via cuFFT on each slab").  Plans use 64-bit sizes
// (cufftMakePlanMany64) so nmesh >= 2048 (> 2^31 elements) works, and a caller-provided work
// area so all large device memory stays owned by the host framework's allocator.
#include <cufft.h>

#include "abk_common.cuh"

struct abk_fft_plan {
    cufftHandle fwd;
    cufftHandle inv;
    bool has_inv;
    size_t work_bytes;
    int kind;  // 0: 3-D r2c/c2r, 1: batched 2-D r2c, 2: batched 1-D c2c along x
    long long nx, ny, nz;
    size_t inv_work_bytes;
};

#define ABK_CHECK_CUFFT(expr)                                                        \
    do {                                                                             \
        cufftResult _r = (expr);                                                     \
        if (_r != CUFFT_SUCCESS) {                                                   \
            abk_set_error("%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #expr, (int)_r); \
            return ABK_ERR_CUFFT;                                                    \
        }                                                                            \
    } while (0)

static int make_handle(cufftHandle *h, int rank, long long *n, long long *inembed, long long istride, long long idist,
                       long long *onembed, long long ostride, long long odist, cufftType type, long long batch,
                       size_t *work)
{
    ABK_CHECK_CUFFT(cufftCreate(h));
    ABK_CHECK_CUFFT(cufftSetAutoAllocation(*h, 0));
    ABK_CHECK_CUFFT(cufftMakePlanMany64(*h, rank, n, inembed, istride, idist, onembed, ostride, odist, type, batch, work));
    return ABK_OK;
}

extern "C" int abk_rfft3_plan_create(abk_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, abk_fft_plan **plan,
                                     size_t *work_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && plan && work_bytes && nx > 0 && ny > 0 && nz > 0, "abk_rfft3_plan_create: bad arguments");
    ABK_CHECK_CUDA(cudaSetDevice(ctx->device));
    abk_fft_plan *p = new abk_fft_plan();
    p->kind = 0;
    long long n[3] = {nx, ny, nz};
    const long long nzc = nz / 2 + 1;
    long long rembed[3] = {nx, ny, 2 * nzc};
    long long cembed[3] = {nx, ny, nzc};
    size_t w1 = 0;
    int rc = make_handle(&p->fwd, 3, n, rembed, 1, nx * ny * 2 * nzc, cembed, 1, nx * ny * nzc, CUFFT_R2C, 1, &w1);
    if (rc) { delete p; return rc; }
    p->has_inv = false;  // the C2R plan (xi(r) path) is created on first use
    p->nx = nx; p->ny = ny; p->nz = nz;
    p->inv_work_bytes = 0;
    p->work_bytes = w1;
    *work_bytes = p->work_bytes;
    *plan = p;
    return ABK_OK;
}

extern "C" int abk_fft_yz_plan_create(abk_ctx *ctx, int64_t nplanes, int64_t ny, int64_t nz, abk_fft_plan **plan,
                                      size_t *work_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && plan && work_bytes && nplanes > 0 && ny > 0 && nz > 0, "abk_fft_yz_plan_create: bad arguments");
    ABK_CHECK_CUDA(cudaSetDevice(ctx->device));
    abk_fft_plan *p = new abk_fft_plan();
    p->kind = 1;
    long long n[2] = {ny, nz};
    const long long nzc = nz / 2 + 1;
    long long rembed[2] = {ny, 2 * nzc};
    long long cembed[2] = {ny, nzc};
    size_t w1 = 0;
    int rc = make_handle(&p->fwd, 2, n, rembed, 1, ny * 2 * nzc, cembed, 1, ny * nzc, CUFFT_R2C, nplanes, &w1);
    if (rc) { delete p; return rc; }
    p->has_inv = false;
    p->work_bytes = w1;
    *work_bytes = w1;
    *plan = p;
    return ABK_OK;
}

extern "C" int abk_fft_x_plan_create(abk_ctx *ctx, int64_t nx, int64_t nrows, int64_t nzc, abk_fft_plan **plan,
                                     size_t *work_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && plan && work_bytes && nx > 0 && nrows > 0 && nzc > 0, "abk_fft_x_plan_create: bad arguments");
    ABK_CHECK_CUDA(cudaSetDevice(ctx->device));
    abk_fft_plan *p = new abk_fft_plan();
    p->kind = 2;
    // layout [x][row][nzc]: transform along x for each of nrows*nzc columns: stride = nrows*nzc, dist = 1
    long long n[1] = {nx};
    long long embed[1] = {nx};
    const long long cols = nrows * nzc;
    size_t w1 = 0;
    int rc = make_handle(&p->fwd, 1, n, embed, cols, 1, embed, cols, 1, CUFFT_C2C, cols, &w1);
    if (rc) { delete p; return rc; }
    p->has_inv = false;
    p->work_bytes = w1;
    *work_bytes = w1;
    *plan = p;
    return ABK_OK;
}

static int prep(abk_ctx *ctx, abk_fft_plan *plan, cufftHandle h, void *work, size_t work_bytes)
{
    ABK_REQUIRE(ctx && plan, "fft exec: null argument");
    if (work_bytes < plan->work_bytes) {
        abk_set_error("fft exec: work area %zu < %zu", work_bytes, plan->work_bytes);
        return ABK_ERR_SCRATCH;
    }
    ABK_CHECK_CUFFT(cufftSetStream(h, ctx->stream));
    if (plan->work_bytes) ABK_CHECK_CUFFT(cufftSetWorkArea(h, work));
    return ABK_OK;
}

extern "C" int abk_rfft3_exec(abk_ctx *ctx, abk_fft_plan *plan, float *grid, void *work, size_t work_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && plan && plan->kind == 0, "abk_rfft3_exec: not a 3-D plan");
    int rc = prep(ctx, plan, plan->fwd, work, work_bytes);
    if (rc) return rc;
    if (ctx->prof_on) abk_prof_begin(ctx, ABK_K_FFT);
    ABK_CHECK_CUFFT(cufftExecR2C(plan->fwd, (cufftReal *)grid, (cufftComplex *)grid));
    if (ctx->prof_on) abk_prof_end(ctx);
    ctx->launches += 1;  // cuFFT launches several kernels; counted as one library call
    return ABK_OK;
}

extern "C" int abk_irfft3_exec(abk_ctx *ctx, abk_fft_plan *plan, float *grid, void *work, size_t work_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && plan && plan->kind == 0, "abk_irfft3_exec: not a 3-D plan");
    if (!plan->has_inv) {
        ABK_CHECK_CUDA(cudaSetDevice(ctx->device));
        long long n[3] = {plan->nx, plan->ny, plan->nz};
        const long long nzc = plan->nz / 2 + 1;
        long long rembed[3] = {plan->nx, plan->ny, 2 * nzc};
        long long cembed[3] = {plan->nx, plan->ny, nzc};
        int rc = make_handle(&plan->inv, 3, n, cembed, 1, plan->nx * plan->ny * nzc, rembed, 1,
                             plan->nx * plan->ny * 2 * nzc, CUFFT_C2R, 1, &plan->inv_work_bytes);
        if (rc) return rc;
        plan->has_inv = true;
    }
    if (work_bytes < plan->inv_work_bytes) {
        abk_set_error("abk_irfft3_exec: work area %zu < %zu", work_bytes, plan->inv_work_bytes);
        return ABK_ERR_SCRATCH;
    }
    ABK_CHECK_CUFFT(cufftSetStream(plan->inv, ctx->stream));
    if (plan->inv_work_bytes) ABK_CHECK_CUFFT(cufftSetWorkArea(plan->inv, work));
    ABK_CHECK_CUFFT(cufftExecC2R(plan->inv, (cufftComplex *)grid, (cufftReal *)grid));
    ctx->launches += 1;
    return ABK_OK;
}

extern "C" int abk_fft_exec_generic(abk_ctx *ctx, abk_fft_plan *plan, void *data, void *work, size_t work_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && plan, "abk_fft_exec_generic: null argument");
    int rc = prep(ctx, plan, plan->fwd, work, work_bytes);
    if (rc) return rc;
    if (ctx->prof_on) abk_prof_begin(ctx, ABK_K_FFT);
    if (plan->kind == 2) ABK_CHECK_CUFFT(cufftExecC2C(plan->fwd, (cufftComplex *)data, (cufftComplex *)data, CUFFT_FORWARD));
    else ABK_CHECK_CUFFT(cufftExecR2C(plan->fwd, (cufftReal *)data, (cufftComplex *)data));
    if (ctx->prof_on) abk_prof_end(ctx);
    ctx->launches += 1;
    return ABK_OK;
}

extern "C" int abk_fft_plan_destroy(abk_fft_plan *plan)
{
    if (!plan) return ABK_OK;
    cufftDestroy(plan->fwd);
    if (plan->has_inv) cufftDestroy(plan->inv);
    delete plan;
    return ABK_OK;
}

// In-place D2Z transform of a padded float64 grid [nx][ny][2 (nz/2+1)] -> complex128 [nx][ny][nz/2+1] (numpy rfftn
// conventions, unnormalised).  The plan is created and destroyed per call (this is not the fast path); cuFFT allocates its
// own work area here.
extern "C" int abk_rfft3_f64(abk_ctx *ctx, double *grid_inplace, int64_t nx, int64_t ny, int64_t nz)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && grid_inplace && nx > 0 && ny > 0 && nz > 0, "abk_rfft3_f64: bad arguments");
    cufftHandle h;
    long long n[3] = {nx, ny, nz};
    const long long nzc = nz / 2 + 1;
    long long rembed[3] = {nx, ny, 2 * nzc}, cembed[3] = {nx, ny, nzc};
    size_t work = 0;
    cufftResult r = cufftCreate(&h);
    if (r == CUFFT_SUCCESS) r = cufftMakePlanMany64(h, 3, n, rembed, 1, nx * ny * 2 * nzc, cembed, 1, nx * ny * nzc, CUFFT_D2Z, 1, &work);
    if (r == CUFFT_SUCCESS) r = cufftSetStream(h, ctx->stream);
    if (ctx->prof_on) abk_prof_begin(ctx, ABK_K_FFT);
    if (r == CUFFT_SUCCESS) r = cufftExecD2Z(h, grid_inplace, (cufftDoubleComplex *)grid_inplace);
    if (ctx->prof_on) abk_prof_end(ctx);
    ctx->launches++;
    cudaStreamSynchronize(ctx->stream);  // the plan (and its work area) is destroyed below
    cufftDestroy(h);
    if (r != CUFFT_SUCCESS) {
        abk_set_error("abk_rfft3_f64: cufft error %d", (int)r);
        return ABK_ERR_CUFFT;
    }
    return ABK_OK;
}

