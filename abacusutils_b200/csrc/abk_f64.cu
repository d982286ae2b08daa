// float64 support of the path (reference: tsc.py:155-165, :394-507; power_spectrum.py:1040-1078).
//
// The reference's TSC kernel computes in the dtype of the POSITIONS (ftype = positions.dtype.type, tsc.py:400) and
// accumulates into whatever dtype the grid has; calc_power(dtype=float64) honours the dtype on its non-interlaced branch
// only (field, rfftn and the 1/n^3 + window division in float64; bin_kmu always sums in float32).  float32 is the fast
// path of this library (tile bucketing + walk kernel); any other combination of position / grid dtype takes the kernels
// here: one thread per particle, the 27 stencil points added with global reductions in the grid's dtype, arithmetic in the
// positions' dtype -- the reference's formulation, typed.  It is correct to float64 round-off and HBM/atomic-bound, not
// tuned: the reference itself warns that float32 is the performance path (tsc.py:157-160).
#include "abk_common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ T wrap_coord_t(T v, double box)
{
    if ((double)v >= box) return (T)((double)v - box);
    if (v < (T)0) return (T)((double)v + box);
    return v;
}

template <typename T> __device__ __forceinline__ T rint_t(T x);
template <> __device__ __forceinline__ float rint_t<float>(float x) { return rintf(x); }
template <> __device__ __forceinline__ double rint_t<double>(double x) { return rint(x); }

// tsc.py:408-507 with ftype = P (positions), grid dtype G; cic = 1: analysis/cic.py:29-67 (always evaluated in double there)
template <typename P, typename G, typename WT>
__global__ void __launch_bounds__(256) tsc_typed_kernel(const P *__restrict__ pos, const WT *__restrict__ w, int64_t N, G *__restrict__ grid, int nx,
                                                        int ny, int nz, int64_t ldz, double box, double offset, int wrap, int cic, double wscale)
{
    const P inv_hx = (P)(nx / box), inv_hy = (P)(ny / box), inv_hz = (P)(nz / box);
    const P off = (P)offset;
    const int64_t sx = (int64_t)ny * ldz;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        P c[3] = {pos[3 * n], pos[3 * n + 1], pos[3 * n + 2]};
        const P W = (P)((w ? (double)w[n] : 1.0) * wscale);
        int cell[3];
        P wt[3][3];
        const int dim[3] = {nx, ny, nz};
        const P inv_h[3] = {inv_hx, inv_hy, inv_hz};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (wrap) c[a] = wrap_coord_t<P>(c[a], box);
            if (cic) {
                const double p = ((double)(c[a] + off) / box) * (double)dim[a];
                const double r = rint(p);
                const P d = (P)(r - p);
                cell[a] = (int)r;
                wt[a][0] = d > (P)0 ? d : (P)0;
                wt[a][1] = (P)1 - (d < (P)0 ? -d : d);
                wt[a][2] = d < (P)0 ? -d : (P)0;
            } else {
                const P p = (c[a] + off) * inv_h[a];
                const P r = rint_t<P>(p);
                const P d = r - p;
                cell[a] = (int)r;
                const P am = (P)0.5 + d, ap = (P)0.5 - d;
                wt[a][0] = (P)0.5 * am * am;
                wt[a][1] = (P)0.75 - d * d;
                wt[a][2] = (P)0.5 * ap * ap;
            }
        }
        for (int a = 0; a < 3; a++) {
            const int64_t gx = abk_wrap_cell(cell[0] + a - 1, nx);
            for (int b = 0; b < 3; b++) {
                const int gy = abk_wrap_cell(cell[1] + b - 1, ny);
                const P wxy = wt[0][a] * wt[1][b];
                for (int k = 0; k < 3; k++) {
                    const int gz = abk_wrap_cell(cell[2] + k - 1, nz);
                    atomicAdd(grid + gx * sx + (int64_t)gy * ldz + gz, (G)(wxy * wt[2][k] * W));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) normalize_f64_kernel(double *__restrict__ grid, int64_t nrows, int64_t nz, int64_t ldz, double norm)
{
    const int64_t total = nrows * ldz;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        if (t % ldz < nz) grid[t] = grid[t] * norm - 1.0;  // power_spectrum.py:893-899
}

// power_spectrum.py:1058-1069 + :707-727 on a complex128 spectrum: f *= inv_size; f /= (W_i W_j) W_k (float32 product, that
// association); power = |f|^2 (or Re(conj(f) f2)), written as the float32 mesh the binning kernels read
__global__ void __launch_bounds__(256) power_from_z_kernel(const double2 *__restrict__ f1, const double2 *__restrict__ f2, const float *__restrict__ W,
                                                           int n, int nzc, double inv_size, float *__restrict__ out)
{
    const int64_t total = (int64_t)n * n * nzc;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % nzc), j = (int)((t / nzc) % n), i = (int)(t / ((int64_t)nzc * n));
        double s = inv_size;
        if (W) s /= (double)((W[i] * W[j]) * W[k]);
        const double2 a = f1[t];
        double p;
        if (f2) {
            const double2 b = f2[t];
            p = (a.x * b.x + a.y * b.y) * s * s;
        } else {
            p = (a.x * a.x + a.y * a.y) * s * s;
        }
        out[t] = (float)p;
    }
}

int grid_for64(const abk_ctx *ctx, int64_t work, int threads)
{
    int64_t blocks = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)ctx->num_sms * 16;
    return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace

extern "C" int abk_tsc_deposit_typed(abk_ctx *ctx, const void *pos, int pos_f64, const void *w, int w_f64, int64_t N, void *grid,
                                     int grid_f64, int nx, int ny, int nz, int64_t ldz, double box, double offset, int wrap)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && grid && (pos || N == 0) && N >= 0 && nx > 0 && ny > 0 && nz > 0 && ldz >= nz && box > 0,
                "abk_tsc_deposit_typed: bad arguments");
    if (N == 0) return ABK_OK;
    const int blocks = grid_for64(ctx, N, 256);
    const int cic = ctx->scheme == 1;
    const double ws = ctx->wscale;
#define ABK_TYPED(P, G, WT)                                                                                                      \
    ABK_LAUNCH(ctx, ABK_K_NAIVE_DEPOSIT, (tsc_typed_kernel<P, G, WT><<<blocks, 256, 0, ctx->stream>>>(                            \
                                             (const P *)pos, (const WT *)w, N, (G *)grid, nx, ny, nz, ldz, box, offset, wrap, cic, ws)))
    if (pos_f64) {
        if (grid_f64) { if (w_f64) ABK_TYPED(double, double, double); else ABK_TYPED(double, double, float); }
        else { if (w_f64) ABK_TYPED(double, float, double); else ABK_TYPED(double, float, float); }
    } else {
        if (grid_f64) { if (w_f64) ABK_TYPED(float, double, double); else ABK_TYPED(float, double, float); }
        else { if (w_f64) ABK_TYPED(float, float, double); else ABK_TYPED(float, float, float); }
    }
#undef ABK_TYPED
    return ABK_OK;
}

extern "C" int abk_normalize_field_f64(abk_ctx *ctx, double *grid, int64_t nx, int64_t ny, int64_t nz, int64_t ldz, double size_total,
                                       double tot_weight)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && grid && nx > 0 && ny > 0 && nz > 0 && ldz >= nz, "abk_normalize_field_f64: bad arguments");
    ABK_REQUIRE(tot_weight != 0.0, "abk_normalize_field_f64: total weight is zero");
    ABK_LAUNCH(ctx, ABK_K_NORMALIZE, normalize_f64_kernel<<<grid_for64(ctx, nx * ny * ldz, 256), 256, 0, ctx->stream>>>(
                                         grid, nx * ny, nz, ldz, size_total / tot_weight));
    return ABK_OK;
}

extern "C" int abk_power_from_f64(abk_ctx *ctx, const void *f1, const void *f2, const float *W, int n, double inv_size, float *out)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && f1 && out && n > 0, "abk_power_from_f64: bad arguments");
    const int nzc = n / 2 + 1;
    ABK_LAUNCH(ctx, ABK_K_RAW_POWER, power_from_z_kernel<<<grid_for64(ctx, (int64_t)n * n * nzc, 256), 256, 0, ctx->stream>>>(
                                         (const double2 *)f1, (const double2 *)f2, W, n, nzc, inv_size, out));
    return ABK_OK;
}
