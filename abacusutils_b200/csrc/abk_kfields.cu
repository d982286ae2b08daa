// Elementwise k-space field helpers used by the reference's ZCV modules:
//   get_delta_mu2       analysis/power_spectrum.py:577-617
//   get_smoothing       :539-574
//   expand_poles_to_3d  :450-520  (+ linear_interp :523-536, P_n :121-147)
// One warp per (i,j) row of the (n, n, n/2+1) mesh, lanes on k: coalesced stores, no reuse -> HBM-bound.
#include "abk_common.cuh"

namespace {

__device__ __forceinline__ int fold(int i, int n) { return (i < n / 2) ? i : i - n; }

template <typename F>
__device__ __forceinline__ void for_each_mode(int n, int nzc, F f)
{
    const int64_t nrows = (int64_t)n * n;
    const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < nrows;
         row += (int64_t)gridDim.x * warps_per_block) {
        const int i = (int)(row / n), j = (int)(row % n);
        const int ii = fold(i, n), jj = fold(j, n);
        const int ij2 = ii * ii + jj * jj;
        for (int k = lane; k < nzc; k += 32) {
            const float kmag2 = (float)(ij2 + k * k);
            const float mu2 = kmag2 > 0.0f ? __fdiv_rn((float)(k * k), kmag2) : 0.0f;
            f(row * nzc + k, kmag2, mu2);
        }
    }
}

__global__ void __launch_bounds__(256) delta_mu2_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int n, int nzc)
{
    for_each_mode(n, nzc, [&](int64_t idx, float, float mu2) {
        const float2 a = in[idx];
        out[idx] = make_float2(a.x * mu2, a.y * mu2);
    });
}

__global__ void __launch_bounds__(256) smoothing_kernel(float *__restrict__ out, int n, int nzc, float dk2, float R2)
{
    for_each_mode(n, nzc, [&](int64_t idx, float kmag2, float) {
        const float t = (-kmag2 * dk2) * R2;   // float32 product, then /2.0 and exp in double (power_spectrum.py:573)
        out[idx] = (float)exp((double)t / 2.0);
    });
}

struct ExpandArgs {
    const float *k_ell;   // [Nk] uniform grid
    const float *P_ell;   // [Np][Nk]
    const float *coef;    // [Np][ABK_POLE_NCOEF]: P_l(mu) as polynomial in mu (no 2l+1 factor)
    int Nk, Np;
    int ell[ABK_MAX_POLES];
    float dk;
    int even_only;
};

__global__ void __launch_bounds__(256) expand_poles_kernel(float *__restrict__ out, int n, int nzc, ExpandArgs A)
{
    const float x0 = A.k_ell[0], xN = A.k_ell[A.Nk - 1], dx = A.k_ell[1] - A.k_ell[0];
    for_each_mode(n, nzc, [&](int64_t idx, float kmag2, float mu2) {
        const float xd = sqrtf(kmag2) * A.dk;
        // linear_interp on the uniform grid (power_spectrum.py:523-536)
        int fl = 0;
        float frac = 0.0f;
        int mode = 0;  // 0 interior, 1 clamp low, 2 clamp high
        if (xd <= x0) mode = 1;
        else if (xd >= xN) mode = 2;
        else {
            const float f = __fdiv_rn(xd - x0, dx);
            fl = (int)f;
            frac = f - (float)fl;
        }
        const float s = A.even_only ? mu2 : sqrtf(mu2);
        float acc = 0.0f;
        for (int ip = 0; ip < A.Np; ip++) {
            const float *y = A.P_ell + (size_t)ip * A.Nk;
            float yd;
            if (mode == 1) yd = y[0];
            else if (mode == 2) yd = y[A.Nk - 1];
            else yd = y[fl] + frac * (y[fl + 1] - y[fl]);
            if (A.ell[ip] != 0) {
                const float *c = A.coef + ip * ABK_POLE_NCOEF;
                float pw;
                if (A.even_only) {
                    pw = c[10];
                    pw = fmaf(pw, s, c[8]); pw = fmaf(pw, s, c[6]); pw = fmaf(pw, s, c[4]);
                    pw = fmaf(pw, s, c[2]); pw = fmaf(pw, s, c[0]);
                } else {
                    pw = c[10];
                    for (int m = 9; m >= 0; m--) pw = fmaf(pw, s, c[m]);
                }
                yd *= pw;
            }
            acc += yd;
        }
        out[idx] = acc;
    });
}

int blocks_for(const abk_ctx *ctx, int64_t nrows)
{
    int64_t b = (nrows + 7) / 8;
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int abk_delta_mu2(abk_ctx *ctx, const void *delta, void *out, int n)
{
    ABK_REQUIRE(ctx && delta && out && n > 0 && n <= 32767, "abk_delta_mu2: bad arguments");
    ABK_LAUNCH(ctx, ABK_K_MISC, delta_mu2_kernel<<<blocks_for(ctx, (int64_t)n * n), 256, 0, ctx->stream>>>(
                                    (const float2 *)delta, (float2 *)out, n, n / 2 + 1));
    return ABK_OK;
}

extern "C" int abk_smoothing(abk_ctx *ctx, float *out, int n, double L, double R)
{
    ABK_REQUIRE(ctx && out && n > 0 && n <= 32767 && L > 0, "abk_smoothing: bad arguments");
    const float dk = (float)(2.0 * M_PI / L);
    const float dk2 = dk * dk;
    const float R2 = (float)(R * R);
    ABK_LAUNCH(ctx, ABK_K_MISC, smoothing_kernel<<<blocks_for(ctx, (int64_t)n * n), 256, 0, ctx->stream>>>(out, n, n / 2 + 1, dk2, R2));
    return ABK_OK;
}

extern "C" int abk_expand_poles_to_3d(abk_ctx *ctx, float *out, int n, double L, const float *k_ell, const float *P_ell,
                                      int Nk, const int32_t *poles_h, int Np, const float *coef)
{
    ABK_REQUIRE(ctx && out && k_ell && P_ell && coef && poles_h && n > 0 && n <= 32767 && Nk >= 2 && Np >= 1 && Np <= ABK_MAX_POLES,
                "abk_expand_poles_to_3d: bad arguments");
    ExpandArgs A;
    A.k_ell = k_ell; A.P_ell = P_ell; A.coef = coef; A.Nk = Nk; A.Np = Np;
    A.dk = (float)(2.0 * M_PI / L);
    A.even_only = 1;
    for (int p = 0; p < ABK_MAX_POLES; p++) A.ell[p] = 0;
    for (int p = 0; p < Np; p++) {
        ABK_REQUIRE(poles_h[p] >= 0 && poles_h[p] <= 10, "abk_expand_poles_to_3d: pole %d out of range", poles_h[p]);
        A.ell[p] = poles_h[p];
        if (poles_h[p] & 1) A.even_only = 0;
    }
    ABK_LAUNCH(ctx, ABK_K_MISC, expand_poles_kernel<<<blocks_for(ctx, (int64_t)n * n), 256, 0, ctx->stream>>>(out, n, n / 2 + 1, A));
    return ABK_OK;
}
