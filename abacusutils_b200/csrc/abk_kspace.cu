// Real-space normalisation and the fused k-space kernels (sm_100a).
//
// Replaces the reference's Numba / NumPy passes
//   normalize_field     analysis/power_spectrum.py:860-901
//   _normalize          :1073-1078          (x 1/n^3 after the FFT)
//   shift_field_fft     :904-948            (interlacing combination)
//   window compensation :1062-1070          (NumPy broadcast divide, materialises a temporary)
//   get_raw_power       :707-727            (|f|^2 or Re(conj f1 f2), materialises P(k))
//   bin_kmu / P_n       :150-300, :121-147  ((k,mu) wedges + Legendre multipoles)
// The reference makes five passes over n^2(n/2+1)-sized arrays; abk_power_bin reads each complex
// mesh exactly once and writes only the O(Nk*Nmu) bin sums.
#include "abk_common.cuh"

namespace {

__device__ __forceinline__ int fold(int i, int n) { return (i < n / 2) ? i : i - n; }

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) normalize_kernel(float *__restrict__ grid, int64_t nrows, int64_t nz,
                                                        int64_t ldz, float norm)
{
    // one warp per row chunk; rows are ldz apart, only nz entries are valid
    const int64_t total = nrows * ldz;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t z = t % ldz;
        if (z < nz) grid[t] = fmaf(grid[t], norm, -1.0f);
    }
}

__global__ void __launch_bounds__(256) normalize_kernel_v2(float2 *__restrict__ grid, int64_t nrows, int64_t nz,
                                                           int64_t ldz2, float norm)
{
    const int64_t total = nrows * ldz2;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t z = (t % ldz2) * 2;
        if (z < nz) {
            float2 v = grid[t];
            v.x = fmaf(v.x, norm, -1.0f);
            if (z + 1 < nz) v.y = fmaf(v.y, norm, -1.0f);
            grid[t] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// finishing arithmetic shared by abk_field_fft_finish and abk_power_bin
struct FinishArgs {
    const float2 *fs;  // shifted mesh or null
    const float *W;    // window table or null
    float scale;
    float inv_n;
    int n;
};

__device__ __forceinline__ float2 finish_mode(float2 a, const FinishArgs &F, int64_t idx, int ii, int jj, int k,
                                              float wij)
{
    if (F.fs) {
        // power_spectrum.py:935-948: phase = exp(i * 0.5 d * (kx+ky+kz)) = exp(i*pi*(i'+j'+k)/n)
        const float2 b = __ldcs(F.fs + idx);
        float sn, cs;
        sincospif((float)(ii + jj + k) * F.inv_n, &sn, &cs);
        a.x += b.x * cs - b.y * sn;
        a.y += b.x * sn + b.y * cs;
    }
    a.x *= F.scale;
    a.y *= F.scale;
    if (F.W) {
        const float ww = wij * F.W[k];  // (W_i * W_j) * W_k, power_spectrum.py:1065-1069
        a.x = __fdiv_rn(a.x, ww);
        a.y = __fdiv_rn(a.y, ww);
    }
    return a;
}

__global__ void __launch_bounds__(256) finish_kernel(float2 *__restrict__ f, FinishArgs F, abk_kmesh M)
{
    const int nj = M.j1 - M.j0;
    const int64_t nrows = (int64_t)(M.i1 - M.i0) * nj;
    const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < nrows;
         row += (int64_t)gridDim.x * warps_per_block) {
        const int il = (int)(row / nj), jl = (int)(row % nj);
        const int i = M.i0 + il, j = M.j0 + jl;
        const int ii = fold(i, M.n), jj = fold(j, M.n);
        const float wij = F.W ? F.W[i] * F.W[j] : 1.0f;
        const int64_t base = il * M.stride_i + jl * M.stride_j;
        for (int k = lane; k < M.nzc; k += 32) {
            const float2 a = f[base + k];
            f[base + k] = finish_mode(a, F, base + k, ii, jj, k, wij);
        }
    }
}

// ------------------------------------------------------------------------------------------
// (k,mu) binning.  One warp walks one (i,j) row with lanes on consecutive k (coalesced 256-byte
// loads).  Within a row the bin index is monotone in k (power_spectrum.py:166-170), so lanes of
// equal bin form contiguous runs: a segmented shuffle reduction leaves each run's sums in its head
// lane, and the head lanes (distinct bins) update a WARP-PRIVATE shared-memory table with plain
// read-modify-writes: no atomics in the main loop (a float shared atomic is a CAS loop on sm_100a).
// Tables are flushed to the global double / u64 sums once per warp.
struct BinArgs {
    abk_kmesh M;
    const float2 *f1, *f2;
    FinishArgs F1, F2;
    const float *real_in;
    int finish;
    const float *kedges2, *muedges2;
    int Nk, Nmu, Np, Npn;  // Npn: poles with ell != 0
    const float *pole_coef;
    int pole_ell[ABK_MAX_POLES];
    int even_only;
    unsigned long long *counts;
    double *sum_p, *sum_k, *sum_poles;
};

template <typename T>
__device__ __forceinline__ T seg_reduce(T v, int key, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T o = __shfl_down_sync(0xffffffffu, v, off);
        const int ko = __shfl_down_sync(0xffffffffu, key, off);
        if (lane + off < 32 && ko == key) v += o;
    }
    return v;
}

// number of table entries e[1..N] strictly below x  (== np.searchsorted(e[1:], x, 'left'))
__device__ __forceinline__ int count_below(const float *__restrict__ e, int N, float x)
{
    int lo = 0, hi = N;  // answer in [0, N]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (e[1 + mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <bool SMEM_TABLES>
__global__ void __launch_bounds__(512) power_bin_kernel(BinArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Nb = A.Nk * A.Nmu;
    float *s_ke = reinterpret_cast<float *>(smem_raw);
    float *s_me = s_ke + (A.Nk + 1);
    float *s_coef = s_me + (A.Nmu + 1);                     // Npn * NCOEF
    int *s_pidx = reinterpret_cast<int *>(s_coef + A.Npn * ABK_POLE_NCOEF);  // Npn: row in sum_poles
    const int hdr = (A.Nk + 1) + (A.Nmu + 1) + A.Npn * ABK_POLE_NCOEF + A.Npn;
    const int hdr_al = (hdr + 3) & ~3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    for (int t = threadIdx.x; t <= A.Nk; t += blockDim.x) s_ke[t] = A.kedges2[t];
    for (int t = threadIdx.x; t <= A.Nmu; t += blockDim.x) s_me[t] = A.muedges2[t];
    if (threadIdx.x == 0) {
        int q = 0;
        for (int p = 0; p < A.Np; p++)
            if (A.pole_ell[p] != 0) {
                for (int c = 0; c < ABK_POLE_NCOEF; c++) s_coef[q * ABK_POLE_NCOEF + c] = A.pole_coef[p * ABK_POLE_NCOEF + c];
                s_pidx[q] = p;
                q++;
            }
    }
    // warp-private tables
    const int tbl_words = 3 * Nb + A.Npn * A.Nk;
    uint32_t *t_cnt = nullptr;
    float *t_p = nullptr, *t_k = nullptr, *t_pl = nullptr;
    if (SMEM_TABLES) {
        uint32_t *base = reinterpret_cast<uint32_t *>(smem_raw) + hdr_al + (size_t)warp * tbl_words;
        t_cnt = base;
        t_p = reinterpret_cast<float *>(base + Nb);
        t_k = t_p + Nb;
        t_pl = t_k + Nb;
        for (int t = lane; t < tbl_words; t += 32) base[t] = 0u;
    }
    __syncthreads();

    const abk_kmesh M = A.M;
    const int nj = M.j1 - M.j0;
    const int64_t nrows = (int64_t)(M.i1 - M.i0) * nj;
    const float e_lo = s_ke[0], e_hi = s_ke[A.Nk];

    for (int64_t row = (int64_t)blockIdx.x * nwarps + warp; row < nrows; row += (int64_t)gridDim.x * nwarps) {
        const int il = (int)(row / nj), jl = (int)(row % nj);
        const int i = M.i0 + il, j = M.j0 + jl;
        const int ii = fold(i, M.n), jj = fold(j, M.n);
        const int ij2 = ii * ii + jj * jj;
        const float wij = (A.finish && A.F1.W) ? A.F1.W[i] * A.F1.W[j] : 1.0f;
        const int64_t base = il * M.stride_i + jl * M.stride_j;
        // the row ends where kmag2 >= last edge (power_spectrum.py:249-250): skip the tail chunks
        for (int k0 = 0; k0 < M.nzc; k0 += 32) {
            if ((float)(ij2 + k0 * k0) >= e_hi) break;  // warp-uniform
            const int k = k0 + lane;
            int key = -1, bk = -1;
            float val = 0.0f, kmag2 = 0.0f, mu2 = 0.0f;
            if (k < M.nzc) {
                kmag2 = (float)(ij2 + k * k);
                if (kmag2 >= e_lo && kmag2 < e_hi) {
                    if (A.real_in) {
                        val = __ldcs(A.real_in + base + k);
                    } else {
                        float2 a = __ldcs(A.f1 + base + k);
                        if (A.finish) a = finish_mode(a, A.F1, base + k, ii, jj, k, wij);
                        if (A.f2) {
                            float2 b = __ldcs(A.f2 + base + k);
                            if (A.finish) b = finish_mode(b, A.F2, base + k, ii, jj, k, wij);
                            val = a.x * b.x + a.y * b.y;
                        } else {
                            val = a.x * a.x + a.y * a.y;
                        }
                    }
                    mu2 = kmag2 > 0.0f ? __fdiv_rn((float)(k * k), kmag2) : 0.0f;
                    bk = count_below(s_ke, A.Nk, kmag2);
                    int bmu = count_below(s_me, A.Nmu, mu2);
                    if (bmu > A.Nmu - 1) bmu = A.Nmu - 1;
                    key = bk * A.Nmu + bmu;
                }
            }
            if (__ballot_sync(0xffffffffu, key >= 0) == 0u) continue;
            const float mult = (k == 0) ? 1.0f : 2.0f;
            const float pv = mult * val;
            const uint32_t c_run = seg_reduce<uint32_t>(key >= 0 ? (k == 0 ? 1u : 2u) : 0u, key, lane);
            const float p_run = seg_reduce<float>(pv, key, lane);
            const float k_run = seg_reduce<float>(mult * sqrtf(kmag2), key, lane);
            const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
            const bool head = (key >= 0) && (lane == 0 || key_prev != key);
            if (head) {
                if (SMEM_TABLES) {
                    t_cnt[key] += c_run;
                    t_p[key] += p_run;
                    t_k[key] += k_run;
                } else {
                    atomicAdd(A.counts + key, (unsigned long long)c_run);
                    atomicAdd(A.sum_p + key, (double)p_run);
                    atomicAdd(A.sum_k + key, (double)k_run);
                }
            }
            if (A.Npn > 0) {
                const int bk_prev = __shfl_up_sync(0xffffffffu, bk, 1);
                const bool head_k = (bk >= 0) && (lane == 0 || bk_prev != bk);
                const float s = A.even_only ? mu2 : sqrtf(mu2);
                for (int q = 0; q < A.Npn; q++) {
                    const float *c = s_coef + q * ABK_POLE_NCOEF;
                    float pw;
                    if (A.even_only) {  // polynomial in x = mu^2: coefficients of s^0, s^2, ... s^10
                        pw = c[10];
                        pw = fmaf(pw, s, c[8]); pw = fmaf(pw, s, c[6]); pw = fmaf(pw, s, c[4]);
                        pw = fmaf(pw, s, c[2]); pw = fmaf(pw, s, c[0]);
                    } else {
                        pw = c[10];
#pragma unroll
                        for (int m = 9; m >= 0; m--) pw = fmaf(pw, s, c[m]);
                    }
                    const float pl_run = seg_reduce<float>(bk >= 0 ? pv * pw : 0.0f, bk, lane);
                    if (head_k) {
                        if (SMEM_TABLES) t_pl[q * A.Nk + bk] += pl_run;
                        else atomicAdd(A.sum_poles + (size_t)s_pidx[q] * A.Nk + bk, (double)pl_run);
                    }
                }
            }
        }
    }

    if (SMEM_TABLES) {
        __syncwarp();
        for (int t = lane; t < Nb; t += 32) {
            const uint32_t c = t_cnt[t];
            if (c) {
                atomicAdd(A.counts + t, (unsigned long long)c);
                atomicAdd(A.sum_p + t, (double)t_p[t]);
                atomicAdd(A.sum_k + t, (double)t_k[t]);
            }
        }
        for (int t = lane; t < A.Npn * A.Nk; t += 32) {
            const float v = t_pl[t];
            if (v != 0.0f) atomicAdd(A.sum_poles + (size_t)s_pidx[t / A.Nk] * A.Nk + (t % A.Nk), (double)v);
        }
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_planes_kernel(float *__restrict__ dst, const float *__restrict__ src,
                                                         int64_t nrows, int64_t nz, int64_t ldz)
{
    const int64_t total = nrows * ldz;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        if (t % ldz < nz) dst[t] += src[t];
    }
}

// slab [nxl][ny][nzc] -> per-destination blocks [nxl][nyl_r][nzc] laid out back to back
constexpr int MAX_RANKS = 64;
struct SplitTable { int64_t v[MAX_RANKS + 1]; };

__global__ void __launch_bounds__(256) transpose_pack_kernel(const float2 *__restrict__ slab, float2 *__restrict__ sendbuf,
                                                             int64_t nxl, int64_t ny, int64_t nzc, int nranks,
                                                             SplitTable js)
{
    __shared__ int64_t s_js[MAX_RANKS + 1];
    for (int t = threadIdx.x; t <= nranks; t += blockDim.x) s_js[t] = js.v[t];
    __syncthreads();
    const int64_t nrows = nxl * ny;
    const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < nrows;
         row += (int64_t)gridDim.x * warps_per_block) {
        const int64_t x = row / ny, j = row % ny;
        int r = 0;
        while (j >= s_js[r + 1]) r++;
        const int64_t nyl = s_js[r + 1] - s_js[r];
        const float2 *src = slab + row * nzc;
        float2 *dst = sendbuf + (nxl * s_js[r] + x * nyl + (j - s_js[r])) * nzc;
        for (int64_t k = lane; k < nzc; k += 32) dst[k] = src[k];
    }
}

// power_spectrum.py:707-727 get_raw_power: |f|^2 or Re(conj(f1) f2), materialised
__global__ void __launch_bounds__(256) raw_power_kernel(const float2 *__restrict__ f1, const float2 *__restrict__ f2,
                                                        float *__restrict__ out, int64_t size)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < size; t += (int64_t)gridDim.x * blockDim.x) {
        const float2 a = f1[t];
        if (f2) {
            const float2 b = f2[t];
            out[t] = a.x * b.x + a.y * b.y;
        } else {
            out[t] = a.x * a.x + a.y * a.y;
        }
    }
}

int grid_for(const abk_ctx *ctx, int64_t work_items, int threads, int per_sm)
{
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)ctx->num_sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int check_mesh(const abk_kmesh &M)
{
    ABK_REQUIRE(M.n > 0 && M.nzc > 0 && M.nzc <= M.n / 2 + 1, "bad mesh n=%d nzc=%d", M.n, M.nzc);
    ABK_REQUIRE(0 <= M.i0 && M.i0 <= M.i1 && M.i1 <= M.n && 0 <= M.j0 && M.j0 <= M.j1 && M.j1 <= M.n,
                "bad mesh ranges i[%d,%d) j[%d,%d) n=%d", M.i0, M.i1, M.j0, M.j1, M.n);
    ABK_REQUIRE(M.n <= 32767, "mesh size %d exceeds 32767", M.n);
    return ABK_OK;
}

}  // namespace

// ==============================================================================================
extern "C" int abk_normalize_field(abk_ctx *ctx, float *grid, int64_t nx, int64_t ny, int64_t nz, int64_t ldz,
                                   double size_total, double tot_weight)
{
    ABK_REQUIRE(ctx && grid && nx > 0 && ny > 0 && nz > 0 && ldz >= nz, "abk_normalize_field: bad arguments");
    ABK_REQUIRE(tot_weight != 0.0, "abk_normalize_field: total weight is zero");
    const float norm = (float)(size_total / tot_weight);  // power_spectrum.py:893
    const int64_t nrows = nx * ny;
    if ((ldz % 2 == 0) && (((uintptr_t)grid & 7) == 0)) {
        ABK_LAUNCH(ctx, ABK_K_NORMALIZE,
                   normalize_kernel_v2<<<grid_for(ctx, nrows * (ldz / 2), 256, 16), 256, 0, ctx->stream>>>(
                       (float2 *)grid, nrows, nz, ldz / 2, norm));
    } else {
        ABK_LAUNCH(ctx, ABK_K_NORMALIZE,
                   normalize_kernel<<<grid_for(ctx, nrows * ldz, 256, 16), 256, 0, ctx->stream>>>(grid, nrows, nz, ldz, norm));
    }
    return ABK_OK;
}

extern "C" int abk_field_fft_finish(abk_ctx *ctx, const abk_kmesh *mesh_h, void *f, const void *fs, const float *W,
                                    float scale)
{
    ABK_REQUIRE(ctx && mesh_h && f, "abk_field_fft_finish: null argument");
    int rc = check_mesh(*mesh_h);
    if (rc) return rc;
    FinishArgs F;
    F.fs = (const float2 *)fs;
    F.W = W;
    F.scale = scale;
    F.n = mesh_h->n;
    F.inv_n = 1.0f / (float)mesh_h->n;
    const int64_t nrows = (int64_t)(mesh_h->i1 - mesh_h->i0) * (mesh_h->j1 - mesh_h->j0);
    if (nrows == 0) return ABK_OK;
    ABK_LAUNCH(ctx, ABK_K_FINISH,
               finish_kernel<<<grid_for(ctx, nrows * 32, 256, 8), 256, 0, ctx->stream>>>((float2 *)f, F, *mesh_h));
    return ABK_OK;
}

extern "C" int abk_raw_power(abk_ctx *ctx, const void *f1, const void *f2, float *out, int64_t size)
{
    ABK_REQUIRE(ctx && f1 && out && size >= 0, "abk_raw_power: bad arguments");
    if (size == 0) return ABK_OK;
    ABK_LAUNCH(ctx, ABK_K_RAW_POWER, raw_power_kernel<<<grid_for(ctx, size, 256, 16), 256, 0, ctx->stream>>>((const float2 *)f1, (const float2 *)f2, out, size));
    return ABK_OK;
}

extern "C" int abk_power_bin(abk_ctx *ctx, const abk_bin_request *R)
{
    ABK_REQUIRE(ctx && R, "abk_power_bin: null argument");
    int rc = check_mesh(R->mesh);
    if (rc) return rc;
    ABK_REQUIRE((R->f1 != nullptr) != (R->real_in != nullptr), "abk_power_bin: exactly one of f1 / real_in must be given");
    ABK_REQUIRE(R->Nk >= 1 && R->Nmu >= 1 && R->Np >= 0 && R->Np <= ABK_MAX_POLES, "abk_power_bin: bad bin counts");
    ABK_REQUIRE(R->kedges2 && R->muedges2 && R->counts && R->sum_p && R->sum_k, "abk_power_bin: null table/output");
    ABK_REQUIRE(R->Np == 0 || (R->pole_coef && R->sum_poles), "abk_power_bin: poles need coef/out");
    ABK_REQUIRE(!(R->real_in && (R->f2 || R->finish)), "abk_power_bin: real input excludes f2/finish");
    ABK_REQUIRE(!(R->f2s && !R->f2) && !(R->f1s && !R->f1), "abk_power_bin: shifted mesh without its base mesh");
    ABK_REQUIRE((int64_t)R->Nk * R->Nmu < ((int64_t)1 << 28), "abk_power_bin: too many bins");

    BinArgs A;
    A.M = R->mesh;
    A.f1 = (const float2 *)R->f1;
    A.f2 = (const float2 *)R->f2;
    A.real_in = R->real_in;
    A.finish = R->finish;
    const float inv_n = 1.0f / (float)R->mesh.n;
    A.F1 = FinishArgs{(const float2 *)R->f1s, R->W, R->scale, inv_n, R->mesh.n};
    A.F2 = FinishArgs{(const float2 *)R->f2s, R->W, R->scale, inv_n, R->mesh.n};
    A.kedges2 = R->kedges2;
    A.muedges2 = R->muedges2;
    A.Nk = R->Nk; A.Nmu = R->Nmu; A.Np = R->Np;
    A.pole_coef = R->pole_coef;
    A.counts = R->counts;
    A.sum_p = R->sum_p; A.sum_k = R->sum_k; A.sum_poles = R->sum_poles;

    A.Npn = 0;
    A.even_only = 1;
    for (int p = 0; p < ABK_MAX_POLES; p++) A.pole_ell[p] = 0;
    for (int p = 0; p < R->Np; p++) {
        const int ell = R->pole_ell[p];
        ABK_REQUIRE(ell >= 0 && ell <= 10, "abk_power_bin: pole %d out of range [0,10]", ell);
        A.pole_ell[p] = ell;
        if (ell != 0) A.Npn++;
        if (ell & 1) A.even_only = 0;
    }

    const int64_t nrows = (int64_t)(A.M.i1 - A.M.i0) * (A.M.j1 - A.M.j0);
    if (nrows == 0) return ABK_OK;
    const int64_t Nb = (int64_t)A.Nk * A.Nmu;
    const int hdr = (A.Nk + 1) + (A.Nmu + 1) + A.Npn * ABK_POLE_NCOEF + A.Npn;
    const size_t hdr_bytes = (size_t)((hdr + 3) & ~3) * 4;
    const size_t tbl_bytes = (size_t)(3 * Nb + (int64_t)A.Npn * A.Nk) * 4;
    const size_t budget = (size_t)ctx->smem_optin - 1024;
    int warps = 0;
    if (hdr_bytes < budget) warps = (int)((budget - hdr_bytes) / tbl_bytes);
    if (warps > 16) warps = 16;
    ABK_REQUIRE(hdr_bytes + 64 < budget, "abk_power_bin: edge tables (%zu B) do not fit in shared memory", hdr_bytes);
    if (warps >= 4) {
        const size_t smem = hdr_bytes + (size_t)warps * tbl_bytes;
        ABK_CHECK_CUDA(cudaFuncSetAttribute(power_bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t blocks = (nrows + warps - 1) / warps;
        if (blocks > ctx->num_sms) blocks = ctx->num_sms;
        ABK_LAUNCH(ctx, ABK_K_POWER_BIN, power_bin_kernel<true><<<(unsigned)blocks, warps * 32, smem, ctx->stream>>>(A));
    } else {
        ABK_CHECK_CUDA(cudaFuncSetAttribute(power_bin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hdr_bytes));
        int64_t blocks = (nrows + 15) / 16;
        if (blocks > (int64_t)ctx->num_sms * 2) blocks = (int64_t)ctx->num_sms * 2;
        ABK_LAUNCH(ctx, ABK_K_POWER_BIN, power_bin_kernel<false><<<(unsigned)blocks, 512, hdr_bytes, ctx->stream>>>(A));
    }
    return ABK_OK;
}

extern "C" int abk_add_planes(abk_ctx *ctx, float *dst, const float *src, int64_t nplanes, int64_t ny, int64_t nz,
                              int64_t ldz)
{
    ABK_REQUIRE(ctx && dst && src && nplanes >= 0 && ny > 0 && nz > 0 && ldz >= nz, "abk_add_planes: bad arguments");
    if (nplanes == 0) return ABK_OK;
    ABK_LAUNCH(ctx, ABK_K_ADD_PLANES,
               add_planes_kernel<<<grid_for(ctx, nplanes * ny * ldz, 256, 16), 256, 0, ctx->stream>>>(dst, src, nplanes * ny, nz, ldz));
    return ABK_OK;
}

extern "C" int abk_transpose_pack(abk_ctx *ctx, const void *slab, void *sendbuf, int64_t nxl, int64_t ny, int64_t nzc,
                                  int nranks, const int64_t *jsplit_h)
{
    ABK_REQUIRE(ctx && slab && sendbuf && nxl >= 0 && ny > 0 && nzc > 0 && nranks > 0 && nranks <= MAX_RANKS && jsplit_h,
                "abk_transpose_pack: bad arguments");
    ABK_REQUIRE(jsplit_h[0] == 0 && jsplit_h[nranks] == ny, "abk_transpose_pack: jsplit must run from 0 to ny");
    if (nxl == 0) return ABK_OK;
    SplitTable js;
    for (int r = 0; r <= nranks; r++) js.v[r] = jsplit_h[r];
    ABK_LAUNCH(ctx, ABK_K_TRANSPOSE_PACK, transpose_pack_kernel<<<grid_for(ctx, nxl * ny * 32, 256, 8), 256, 0, ctx->stream>>>(
        (const float2 *)slab, (float2 *)sendbuf, nxl, ny, nzc, nranks, js));
    return ABK_OK;
}
