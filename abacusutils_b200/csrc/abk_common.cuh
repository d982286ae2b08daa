// Internal helpers shared by the libabk translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "abk.h"

struct abk_ctx {
    int device;
    cudaStream_t stream;
    int num_sms;
    int smem_optin;  // max dynamic shared memory per block (opt-in)
    int64_t launches;
    int tile_capacity;  // 0 = auto
    // small device scratch owned by the context (work counters, flags)
    unsigned long long *d_scalars;
};

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
// One warp per y-row of the tile, one lane per z-cell, x walked serially with a rolling
// 3-plane register window.  See abk_tsc.cu.
constexpr int ABK_TX = 8;
constexpr int ABK_TY = 8;
constexpr int ABK_TZ = 32;

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
