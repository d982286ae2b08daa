// Elementwise k-space field helpers used by the reference's ZCV modules:
//   get_delta_mu2       analysis/power_spectrum.py:577-617
//   get_smoothing       :539-574
//   expand_poles_to_3d  :450-520  (+ linear_interp :523-536, P_n :121-147)
//   bin_kppi            :303-412
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

// ---- bin_kppi (power_spectrum.py:303-412) ---------------------------------------------------
// The pi bin depends on k only and the k_perp bin on (i,j) only.  A warp owns (i, 32 consecutive k):
// every lane keeps ONE pi bin for the whole task and walks j with coalesced 128-byte loads; the
// k_perp bin, the skip below kedges2[0] and the reference's early `break` are warp-uniform, so runs
// of equal k_perp bin are summed in registers and flushed with one RED per warp (one per lane only
// when the 32 k straddle a pi edge).
template <typename T>
__global__ void __launch_bounds__(256)
bin_kppi_kernel(const T *__restrict__ w, int n, int nzc, int64_t ldz, const double *__restrict__ kedges2, int Nk,
                const double *__restrict__ piedges2, int Npi, int kperp_f32, int nchunk,
                unsigned long long *__restrict__ counts, double *__restrict__ sums)
{
    extern __shared__ double s_ke[];   // kedges2 [Nk+1]
    for (int t = threadIdx.x; t <= Nk; t += blockDim.x) s_ke[t] = kedges2[t];
    __syncthreads();
    const double klo = s_ke[0], khi = s_ke[Nk];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int64_t ntask = (int64_t)n * nchunk;

    for (int64_t task = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); task < ntask; task += (int64_t)gridDim.x * wpb) {
        const int i = (int)(task / nchunk), k = (int)(task % nchunk) * 32 + lane;
        // pi bin of this lane: first b with kz2 <= piedges2[b+1]; kz2 is an exact integer, compared in
        // double like the reference's int-vs-float promotion; dropped when kz2 >= piedges2[Npi] (:388-395)
        int bpi = -1;
        if (k < nzc) {
            const double kz2 = (double)k * (double)k;
            if (kz2 < piedges2[Npi]) {
                int lo = 0, hi = Npi - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (kz2 <= piedges2[mid + 1]) hi = mid; else lo = mid + 1;
                }
                bpi = lo;
            }
        }
        const unsigned act = __ballot_sync(full, bpi >= 0);
        if (act == 0) continue;
        const int b0 = __shfl_sync(full, bpi, __ffs(act) - 1);
        const bool one_bin = __all_sync(full, bpi < 0 || bpi == b0);
        const unsigned mult = (k == 0) ? 1u : 2u;
        // multiplicity summed over the active lanes (only lane 0 of chunk 0 can hold k == 0)
        const unsigned warp_mult = 2u * (unsigned)__popc(act) - ((task % nchunk) == 0 && (act & 1u) ? 1u : 0u);

        const int ii = fold(i, n), i2 = ii * ii;
        // the reference breaks out of the j loop at the first j with kperp2 >= kedges2[-1] (:379-380)
        int jend = n;
        for (int j = 0; j < n; j++) {
            const int jj = fold(j, n), ij2 = i2 + jj * jj;
            const double kp2 = kperp_f32 ? (double)(float)ij2 : (double)ij2;
            if (kp2 >= khi) { jend = j; break; }
        }
        const T *base = w + (int64_t)i * n * ldz + k;

        int cur_bk = -1;
        unsigned long long run = 0;
        double acc = 0.0;
        auto flush = [&]() {
            if (cur_bk >= 0 && run > 0) {
                const int64_t slot = (int64_t)cur_bk * Npi;
                if (one_bin) {
                    double v = acc;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
                    if (lane == 0) {
                        atomicAdd(&counts[slot + b0], run * warp_mult);
                        atomicAdd(&sums[slot + b0], v);
                    }
                } else if (bpi >= 0) {
                    atomicAdd(&counts[slot + bpi], run * mult);
                    atomicAdd(&sums[slot + bpi], acc);
                }
            }
            run = 0;
            acc = 0.0;
        };

        for (int j0 = 0; j0 < jend; j0 += 4) {
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                v[u] = (bpi >= 0 && j0 + u < jend) ? base[(int64_t)(j0 + u) * ldz] : T(0);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + u;
                if (j >= jend) break;
                const int jj = fold(j, n), ij2 = i2 + jj * jj;
                const double kp2 = kperp_f32 ? (double)(float)ij2 : (double)ij2;
                if (kp2 < klo) continue;   // :375-376
                int lo = 0, hi = Nk - 1;   // first b with kp2 <= kedges2[b+1]  (:382-383), kp2 < kedges2[Nk] here
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (kp2 <= s_ke[mid + 1]) hi = mid; else lo = mid + 1;
                }
                if (lo != cur_bk) { flush(); cur_bk = lo; }
                run++;
                acc += (double)v[u] * (double)mult;
            }
        }
        flush();
    }
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
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && delta && out && n > 0 && n <= 32767, "abk_delta_mu2: bad arguments");
    ABK_LAUNCH(ctx, ABK_K_MISC, delta_mu2_kernel<<<blocks_for(ctx, (int64_t)n * n), 256, 0, ctx->stream>>>(
                                    (const float2 *)delta, (float2 *)out, n, n / 2 + 1));
    return ABK_OK;
}

extern "C" int abk_smoothing(abk_ctx *ctx, float *out, int n, double L, double R)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
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
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
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

extern "C" int abk_bin_kppi(abk_ctx *ctx, const void *weights, int weights_f64, int n, int64_t ldz, const double *kedges2,
                            int Nk, const double *piedges2, int Npi, int kperp_f32, unsigned long long *counts,
                            double *sum_w)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && weights && kedges2 && piedges2 && counts && sum_w, "abk_bin_kppi: null argument");
    ABK_REQUIRE(n > 0 && n <= 32767 && ldz >= n / 2 + 1, "abk_bin_kppi: bad mesh (n=%d, ldz=%lld)", n, (long long)ldz);
    ABK_REQUIRE(Nk >= 1 && Npi >= 1 && (int64_t)Nk * Npi < ((int64_t)1 << 28), "abk_bin_kppi: bad bin counts");
    const size_t smem = (size_t)(Nk + 1) * sizeof(double);
    ABK_REQUIRE(smem + 1024 < (size_t)ctx->smem_optin, "abk_bin_kppi: k edges (%zu B) do not fit in shared memory", smem);
    const int nzc = n / 2 + 1, nchunk = (nzc + 31) / 32;
    const int blocks = blocks_for(ctx, (int64_t)n * nchunk);
    if (weights_f64) {
        ABK_CHECK_CUDA(cudaFuncSetAttribute(bin_kppi_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ABK_LAUNCH(ctx, ABK_K_MISC, bin_kppi_kernel<double><<<blocks, 256, smem, ctx->stream>>>(
                                        (const double *)weights, n, nzc, ldz, kedges2, Nk, piedges2, Npi, kperp_f32, nchunk, counts, sum_w));
    } else {
        ABK_CHECK_CUDA(cudaFuncSetAttribute(bin_kppi_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ABK_LAUNCH(ctx, ABK_K_MISC, bin_kppi_kernel<float><<<blocks, 256, smem, ctx->stream>>>(
                                        (const float *)weights, n, nzc, ldz, kedges2, Nk, piedges2, Npi, kperp_f32, nchunk, counts, sum_w));
    }
    return ABK_OK;
}
