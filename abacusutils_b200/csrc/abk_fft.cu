// FFT stage: power_spectrum.py:980,986,1059 call scipy.fft.rfftn (pocketfft); here the same
// unnormalised forward R2C transform runs in place through cuFFT (BASELINE.json north_star:
// "the real-to-complex 3D FFT via cuFFT on each slab").  Plans use 64-bit sizes
// (cufftMakePlanMany64) so nmesh >= 2048 (> 2^31 elements) works, and a caller-provided work
// area so all large device memory stays owned by the host framework's allocator.
#include <cufft.h>
#include <dlfcn.h>
#include <mutex>
#include <string>

#include "abk_common.cuh"

// ---- which cuFFT ---------------------------------------------------------------------------------
// libabk does not link against libcufft: in a process that has imported torch, the soname libcufft.so.11 is already bound to
// the copy bundled with torch (CUDA 12.8), and that is what a linked libabk would silently get.  The toolkit's cuFFT 11.4
// (CUDA 12.9) transforms the 1024^3 grid in place in 5.9 ms with no work area, the bundled 11.3 in 6.6 ms with a 4.3 GB work
// area (scripts/micro/fft_pad_micro.cu, profiles/r2_micro8.log).  So the library is opened by PATH at first use -- a path with
// a slash is matched by file identity, not by soname, and loads next to torch's copy: $ABK_CUFFT, then $CUDA_HOME/lib64 and
// /usr/local/cuda/lib64, then whatever "libcufft.so.11" resolves to (abk_fft_backend reports the choice).
namespace {
struct CufftApi {
    void *handle = nullptr;
    std::string path;
    int version = 0;
    cufftResult (*Create)(cufftHandle *) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
    cufftResult (*SetAutoAllocation)(cufftHandle, int) = nullptr;
    cufftResult (*MakePlanMany64)(cufftHandle, int, long long *, long long *, long long, long long, long long *, long long, long long,
                                  cufftType, long long, size_t *) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*SetWorkArea)(cufftHandle, void *) = nullptr;
    cufftResult (*ExecR2C)(cufftHandle, cufftReal *, cufftComplex *) = nullptr;
    cufftResult (*ExecC2R)(cufftHandle, cufftComplex *, cufftReal *) = nullptr;
    cufftResult (*ExecC2C)(cufftHandle, cufftComplex *, cufftComplex *, int) = nullptr;
    cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal *, cufftDoubleComplex *) = nullptr;
    cufftResult (*GetVersion)(int *) = nullptr;
};
CufftApi g_fft;
std::once_flag g_fft_once;

bool cufft_try(const std::string &path)
{
    void *h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) return false;
    CufftApi a;
    a.handle = h;
    a.path = path;
#define ABK_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name))
    ABK_SYM(Create, "cufftCreate"); ABK_SYM(Destroy, "cufftDestroy"); ABK_SYM(SetAutoAllocation, "cufftSetAutoAllocation");
    ABK_SYM(MakePlanMany64, "cufftMakePlanMany64"); ABK_SYM(SetStream, "cufftSetStream"); ABK_SYM(SetWorkArea, "cufftSetWorkArea");
    ABK_SYM(ExecR2C, "cufftExecR2C"); ABK_SYM(ExecC2R, "cufftExecC2R"); ABK_SYM(ExecC2C, "cufftExecC2C");
    ABK_SYM(ExecD2Z, "cufftExecD2Z"); ABK_SYM(GetVersion, "cufftGetVersion");
#undef ABK_SYM
    if (!a.Create || !a.Destroy || !a.SetAutoAllocation || !a.MakePlanMany64 || !a.SetStream || !a.SetWorkArea || !a.ExecR2C ||
        !a.ExecC2R || !a.ExecC2C || !a.ExecD2Z || !a.GetVersion) {
        dlclose(h);
        return false;
    }
    a.GetVersion(&a.version);
    g_fft = a;
    return true;
}

void cufft_load()
{
    const char *env = getenv("ABK_CUFFT");
    if (env && *env && cufft_try(env)) return;
    const char *home = getenv("CUDA_HOME");
    if (home && *home && cufft_try(std::string(home) + "/lib64/libcufft.so.11")) return;
    if (cufft_try("/usr/local/cuda/lib64/libcufft.so.11")) return;
    cufft_try("libcufft.so.11");
}

// nullptr (and the error string set) if no cuFFT can be loaded
const CufftApi *cufft_api()
{
    std::call_once(g_fft_once, cufft_load);
    if (!g_fft.handle) {
        abk_set_error("cuFFT not found: tried $ABK_CUFFT, $CUDA_HOME/lib64/libcufft.so.11, /usr/local/cuda/lib64/libcufft.so.11, libcufft.so.11");
        return nullptr;
    }
    return &g_fft;
}
}  // namespace

#define ABK_FFT_API(var)                 \
    const CufftApi *var = cufft_api();   \
    if (!var) return ABK_ERR_CUFFT

extern "C" int abk_fft_backend(char *path, int path_len, int *version)
{
    ABK_FFT_API(F);
    if (path && path_len > 0) snprintf(path, (size_t)path_len, "%s", F->path.c_str());
    if (version) *version = F->version;
    return ABK_OK;
}

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
    ABK_FFT_API(F);
    ABK_CHECK_CUFFT(F->Create(h));
    ABK_CHECK_CUFFT(F->SetAutoAllocation(*h, 0));
    ABK_CHECK_CUFFT(F->MakePlanMany64(*h, rank, n, inembed, istride, idist, onembed, ostride, odist, type, batch, work));
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
    ABK_FFT_API(F);
    ABK_CHECK_CUFFT(F->SetStream(h, ctx->stream));
    if (plan->work_bytes) ABK_CHECK_CUFFT(F->SetWorkArea(h, work));
    return ABK_OK;
}

extern "C" int abk_rfft3_exec(abk_ctx *ctx, abk_fft_plan *plan, float *grid, void *work, size_t work_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && plan && plan->kind == 0, "abk_rfft3_exec: not a 3-D plan");
    int rc = prep(ctx, plan, plan->fwd, work, work_bytes);
    if (rc) return rc;
    if (ctx->prof_on) abk_prof_begin(ctx, ABK_K_FFT);
    ABK_CHECK_CUFFT(g_fft.ExecR2C(plan->fwd, (cufftReal *)grid, (cufftComplex *)grid));
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
    ABK_CHECK_CUFFT(g_fft.SetStream(plan->inv, ctx->stream));
    if (plan->inv_work_bytes) ABK_CHECK_CUFFT(g_fft.SetWorkArea(plan->inv, work));
    ABK_CHECK_CUFFT(g_fft.ExecC2R(plan->inv, (cufftComplex *)grid, (cufftReal *)grid));
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
    if (plan->kind == 2) ABK_CHECK_CUFFT(g_fft.ExecC2C(plan->fwd, (cufftComplex *)data, (cufftComplex *)data, CUFFT_FORWARD));
    else ABK_CHECK_CUFFT(g_fft.ExecR2C(plan->fwd, (cufftReal *)data, (cufftComplex *)data));
    if (ctx->prof_on) abk_prof_end(ctx);
    ctx->launches += 1;
    return ABK_OK;
}

extern "C" int abk_fft_plan_destroy(abk_fft_plan *plan)
{
    if (!plan) return ABK_OK;
    if (g_fft.handle) {      // a plan can only exist if the library was loaded
        g_fft.Destroy(plan->fwd);
        if (plan->has_inv) g_fft.Destroy(plan->inv);
    }
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
    ABK_FFT_API(F);
    cufftHandle h;
    long long n[3] = {nx, ny, nz};
    const long long nzc = nz / 2 + 1;
    long long rembed[3] = {nx, ny, 2 * nzc}, cembed[3] = {nx, ny, nzc};
    size_t work = 0;
    cufftResult r = F->Create(&h);
    if (r == CUFFT_SUCCESS) r = F->MakePlanMany64(h, 3, n, rembed, 1, nx * ny * 2 * nzc, cembed, 1, nx * ny * nzc, CUFFT_D2Z, 1, &work);
    if (r == CUFFT_SUCCESS) r = F->SetStream(h, ctx->stream);
    if (ctx->prof_on) abk_prof_begin(ctx, ABK_K_FFT);
    if (r == CUFFT_SUCCESS) r = F->ExecD2Z(h, grid_inplace, (cufftDoubleComplex *)grid_inplace);
    if (ctx->prof_on) abk_prof_end(ctx);
    ctx->launches++;
    cudaStreamSynchronize(ctx->stream);  // the plan (and its work area) is destroyed below
    F->Destroy(h);
    if (r != CUFFT_SUCCESS) {
        abk_set_error("abk_rfft3_f64: cufft error %d", (int)r);
        return ABK_ERR_CUFFT;
    }
    return ABK_OK;
}

