// Internal helpers shared by the libabk translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "abk.h"

// kernel ids for the optional per-kernel CUDA-event timing (abk_ctx_profile_*)
enum abk_kernel_id {
    ABK_K_WRAP = 0,
    ABK_K_PART_HIST,
    ABK_K_PART_SCATTER,
    ABK_K_SCAN,
    ABK_K_BUCKET_HIST,
    ABK_K_BUCKET_SCATTER,
    ABK_K_TILE_DEPOSIT,
    ABK_K_NAIVE_DEPOSIT,
    ABK_K_NORMALIZE,
    ABK_K_FFT,
    ABK_K_FINISH,
    ABK_K_RAW_POWER,
    ABK_K_POWER_BIN,
    ABK_K_ADD_PLANES,
    ABK_K_TRANSPOSE_PACK,
    ABK_K_MISC,
    ABK_K_COUNT
};

struct abk_prof_rec {
    int id;
    cudaEvent_t a, b;
};

struct abk_ctx {
    int device;
    cudaStream_t stream;
    int num_sms;
    int smem_optin;  // max dynamic shared memory per block (opt-in)
    int64_t launches;
    int tile_capacity;  // 0 = auto
    int scheme;         // mass-assignment scheme of the deposit entry points: 0 TSC, 1 CIC
    int bin_no_sym;     // experiments: force the one-mode-at-a-time binning kernel
    float wscale;       // factor applied to every particle weight when records are written (abk_ctx_set_weight_scale)
    // small device scratch owned by the context (work counters, flags)
    unsigned long long *d_scalars;
    // profiling state
    int prof_on;
    abk_prof_rec *prof_recs;
    int prof_n, prof_cap;
    cudaEvent_t *prof_pool;
    int prof_pool_n, prof_pool_cap;
};

void abk_prof_begin(abk_ctx *ctx, int id);
void abk_prof_end(abk_ctx *ctx);

// Makes the context's device current for the lifetime of the guard and restores the caller's device afterwards, so a
// context of cuda:1 can be used while cuda:0 is current (its streams and buffers belong to device 1).
struct abk_device_guard {
    int prev = -1, dev;
    explicit abk_device_guard(int device) : dev(device)
    {
        if (dev < 0) return;  // null context: the entry point reports the error itself
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~abk_device_guard()
    {
        if (dev >= 0 && prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};

// launch wrapper: device guard, optional event pair around the launch, launch counter, error check
#define ABK_LAUNCH(ctx, id, ...)                 \
    do {                                         \
        abk_device_guard _guard((ctx)->device);  \
        if ((ctx)->prof_on) abk_prof_begin((ctx), (id)); \
        __VA_ARGS__;                             \
        if ((ctx)->prof_on) abk_prof_end((ctx)); \
        ABK_CHECK_LAUNCH(ctx);                   \
    } while (0)

void abk_set_error(const char *fmt, ...);

#define ABK_CHECK_CUDA(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            abk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return ABK_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define ABK_CHECK_LAUNCH(ctx)                                                                 \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            abk_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return ABK_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define ABK_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            abk_set_error(__VA_ARGS__);   \
            return ABK_ERR_INVALID;       \
        }                                 \
    } while (0)

static inline size_t abk_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- deposit tile geometry (cells per tile) -------------------------------------------------
// One warp per y-row of the tile, one lane per z-cell plus one halo lane either side, x walked serially
// with a rolling 3-plane register window.  See abk_tsc.cu.
constexpr int ABK_TX = 8;
constexpr int ABK_TY = 8;
constexpr int ABK_TZ = 30;

struct abk_tile_geom {
    int nx, ny, nz;
    int ntx, nty, ntz;
    int64_t ntiles;
};

static inline abk_tile_geom abk_make_geom(int nx, int ny, int nz)
{
    abk_tile_geom g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.ntx = (nx + ABK_TX - 1) / ABK_TX;
    g.nty = (ny + ABK_TY - 1) / ABK_TY;
    g.ntz = (nz + ABK_TZ - 1) / ABK_TZ;
    g.ntiles = (int64_t)g.ntx * g.nty * g.ntz;
    return g;
}

// exclusive/inclusive scan of uint32 (abk_ctx.cu); `tmp` must hold abk_scan_tmp_bytes(n)
size_t abk_scan_tmp_bytes(int64_t n);
int abk_inclusive_scan_u32(abk_ctx *ctx, uint32_t *data, int64_t n, void *tmp);

#ifdef __CUDACC__
// positive modulo for cell indices that may be slightly outside [0, n)
__device__ __forceinline__ int abk_wrap_cell(int i, int n)
{
    if (i >= n) { i -= n; if (i >= n) i %= n; }
    else if (i < 0) { i += n; if (i < 0) { i %= n; if (i < 0) i += n; } }
    return i;
}
#endif
