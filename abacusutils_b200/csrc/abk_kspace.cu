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

constexpr int MAX_RANKS = 64;
struct SplitTable { int64_t v[MAX_RANKS + 1]; };

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
// (k,mu) binning: arguments shared by the kernels below.
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

// ------------------------------------------------------------------------------------------
// (k,mu) binning kernel.
//
// A warp owns a column task (x-plane i, 32 consecutive k) and walks j.  Along j a lane's |k| moves
// by less than one grid unit per step, so its (k,mu) bin changes only every few modes: each lane
// keeps the running sums of its CURRENT bin in registers (no cross-lane traffic at all) and, when
// the bin changes, sends them to the global float64/u64 sums with fire-and-forget reductions
// (REDG.ADD.F64 / .U64, native on sm_100a).  The sums are replicated NREP times (CTA -> replica)
// to spread same-address contention in L2 and folded by a tiny second kernel.  Bins are tracked
// incrementally (a step changes the bin index by at most a few), so there is no binary search in
// the loop.  Loads stay fully coalesced: lanes hold consecutive k of the same (i,j) row.
constexpr int BIN_NREP = 16;
constexpr int BIN_UNROLL = 4;
constexpr int BIN_PF_DIST = 2;  // prefetch distance (steps) of the symmetric kernel
constexpr int BIN_AJ_SEG = 64;  // symmetric kernel: a task walks this many |j'| (a full column would be up to n/2 + 1 steps)

template <int NPN>
struct LaneAcc {
    int key;  // bk * Nmu + bmu of the running bin, -1 = empty
    int bk;
    unsigned cnt;
    float p, k;
    float pl[NPN > 0 ? NPN : 1];
};

template <int NPN>
__device__ __forceinline__ void lane_flush(const BinArgs &A, const LaneAcc<NPN> &acc, size_t rep_off_bins,
                                           size_t rep_off_poles, const int *s_pidx)
{
    if (acc.key < 0) return;
    atomicAdd(A.counts + rep_off_bins + acc.key, (unsigned long long)acc.cnt);
    atomicAdd(A.sum_p + rep_off_bins + acc.key, (double)acc.p);
    atomicAdd(A.sum_k + rep_off_bins + acc.key, (double)acc.k);
#pragma unroll
    for (int q = 0; q < NPN; q++)
        if (q < A.Npn) atomicAdd(A.sum_poles + rep_off_poles + (size_t)s_pidx[q] * A.Nk + acc.bk, (double)acc.pl[q]);
}

template <int NPN>
__global__ void __launch_bounds__(256) power_bin2_kernel(BinArgs A, unsigned *__restrict__ task_counter, int nrep, int use_tab)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_ke = reinterpret_cast<float *>(smem_raw);
    float *s_me = s_ke + (A.Nk + 1);
    float *s_coef = s_me + (A.Nmu + 1);
    int *s_pidx = reinterpret_cast<int *>(s_coef + A.Npn * ABK_POLE_NCOEF);
    const int lane = threadIdx.x & 31;
    const abk_kmesh M = A.M;
    const int nj = M.j1 - M.j0, ni = M.i1 - M.i0;
    // per-row tables of the fused finish step (use_tab): phasor e^{i pi j' / n} of the interlacing phase and window W[j] for
    // the local j range -- the phase of a mode is the product of a per-lane phasor (i', k) and this one: no sincos per mode
    const int hdr = (A.Nk + 1) + (A.Nmu + 1) + A.Npn * ABK_POLE_NCOEF + A.Npn;
    float2 *s_phj = reinterpret_cast<float2 *>(smem_raw + (size_t)((hdr + 3) & ~3) * 4);
    float *s_Wj = reinterpret_cast<float *>(s_phj + (use_tab ? nj : 0));

    for (int t = threadIdx.x; t <= A.Nk; t += blockDim.x) s_ke[t] = A.kedges2[t];
    for (int t = threadIdx.x; t <= A.Nmu; t += blockDim.x) s_me[t] = A.muedges2[t];
    if (use_tab) {
        for (int t = threadIdx.x; t < nj; t += blockDim.x) {
            float sn, cs;
            sincospif((float)fold(M.j0 + t, M.n) * A.F1.inv_n, &sn, &cs);
            s_phj[t] = make_float2(cs, sn);
            s_Wj[t] = A.F1.W ? A.F1.W[M.j0 + t] : 1.0f;
        }
    }
    if (threadIdx.x == 0) {
        int q = 0;
        for (int p = 0; p < A.Np; p++)
            if (A.pole_ell[p] != 0) {
                for (int c = 0; c < ABK_POLE_NCOEF; c++) s_coef[q * ABK_POLE_NCOEF + c] = A.pole_coef[p * ABK_POLE_NCOEF + c];
                s_pidx[q] = p;
                q++;
            }
    }
    __syncthreads();

    const int Nk = A.Nk, Nmu = A.Nmu;
    const int nchunks = (M.nzc + 31) / 32;
    const unsigned ntasks = (unsigned)ni * nchunks;
    const float e_lo = s_ke[0], e_hi = s_ke[Nk];
    const int rep = (nrep > 1) ? (int)(blockIdx.x % nrep) : 0;
    const size_t rep_off_bins = (size_t)rep * Nk * Nmu, rep_off_poles = (size_t)rep * A.Np * Nk;

    for (;;) {
        unsigned task = 0;
        if (lane == 0) task = atomicAdd(task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= ntasks) break;
        const int il = task / nchunks, k0 = (task % nchunks) * 32;
        const int i = M.i0 + il, ii = fold(i, M.n);
        const int k = k0 + lane;
        const bool k_ok = k < M.nzc;
        const int ik2 = ii * ii + k * k;
        const int ik2_min = ii * ii + k0 * k0;  // smallest |k|^2 of this column at jj = 0
        if ((float)ik2_min >= e_hi) continue;   // the whole column lies beyond the last edge
        const float k2f = (float)(k * k);
        const float Wi = (A.finish && A.F1.W) ? A.F1.W[i] : 1.0f;
        const float Wk = (A.finish && A.F1.W && k_ok) ? A.F1.W[k] : 1.0f;
        const float mult = (k == 0) ? 1.0f : 2.0f;
        const unsigned cmult = (k == 0) ? 1u : 2u;
        float2 e_ik = make_float2(1.0f, 0.0f);  // e^{i pi (i' + k) / n}
        if (use_tab && A.F1.fs) sincospif((float)(ii + k) * A.F1.inv_n, &e_ik.y, &e_ik.x);
        // scale and window are applied to the POWER of the mode (one division), not to the field components
        const float pscale = A.f2 ? A.F1.scale * A.F2.scale : A.F1.scale * A.F1.scale;
        const int wpow = (A.F1.W ? 1 : 0) + ((A.f2 ? A.F2.W : A.F1.W) ? 1 : 0);

        LaneAcc<NPN> acc;
        acc.key = -1; acc.bk = 0; acc.cnt = 0; acc.p = 0.0f; acc.k = 0.0f;
#pragma unroll
        for (int q = 0; q < (NPN > 0 ? NPN : 1); q++) acc.pl[q] = 0.0f;
        int bk = 0, bmu = 0;

        for (int jl0 = 0; jl0 < nj; jl0 += BIN_UNROLL) {
            // ---- issue the loads of BIN_UNROLL rows first ------------------------------------------
            float2 va[BIN_UNROLL], vas[BIN_UNROLL], vb[BIN_UNROLL], vbs[BIN_UNROLL];
            float kmag2[BIN_UNROLL];
            bool use[BIN_UNROLL];
#pragma unroll
            for (int u = 0; u < BIN_UNROLL; u++) {
                const int jl = jl0 + u;
                const int jj = fold(M.j0 + jl, M.n);
                kmag2[u] = (float)(ik2 + jj * jj);
                use[u] = k_ok && jl < nj && kmag2[u] >= e_lo && kmag2[u] < e_hi;
                va[u] = vas[u] = vb[u] = vbs[u] = make_float2(0.0f, 0.0f);
                if (use[u]) {
                    const int64_t idx = il * M.stride_i + jl * M.stride_j + k;
                    if (A.real_in) {
                        va[u].x = __ldcs(A.real_in + idx);
                    } else {
                        va[u] = __ldcs(A.f1 + idx);
                        if (A.finish && A.F1.fs) vas[u] = __ldcs(A.F1.fs + idx);
                        if (A.f2) {
                            vb[u] = __ldcs(A.f2 + idx);
                            if (A.finish && A.F2.fs) vbs[u] = __ldcs(A.F2.fs + idx);
                        }
                    }
                }
            }
            // ---- per-row arithmetic -------------------------------------------------------------------
#pragma unroll
            for (int u = 0; u < BIN_UNROLL; u++) {
                if (!use[u]) continue;
                const int jl = jl0 + u;
                const int j = M.j0 + jl, jj = fold(j, M.n);
                float val;
                if (A.real_in) {
                    val = va[u].x;
                } else {
                    float2 a = va[u], b = vb[u];
                    if (A.finish && use_tab) {
                        const float2 ej = s_phj[jl];
                        const float cs = e_ik.x * ej.x - e_ik.y * ej.y, sn = e_ik.x * ej.y + e_ik.y * ej.x;
                        if (A.F1.fs) {
                            a.x += vas[u].x * cs - vas[u].y * sn;
                            a.y += vas[u].x * sn + vas[u].y * cs;
                        }
                        if (A.f2 && A.F2.fs) {
                            b.x += vbs[u].x * cs - vbs[u].y * sn;
                            b.y += vbs[u].x * sn + vbs[u].y * cs;
                        }
                        float pw = (A.f2 ? (a.x * b.x + a.y * b.y) : (a.x * a.x + a.y * a.y)) * pscale;
                        if (wpow) {
                            const float ww = (Wi * s_Wj[jl]) * Wk;
                            pw = __fdiv_rn(pw, wpow == 2 ? ww * ww : ww);
                        }
                        a = make_float2(pw, 0.0f);
                        b = make_float2(1.0f, 0.0f);
                    } else if (A.finish) {
                        float sn = 0.0f, cs = 1.0f;
                        if (A.F1.fs) {
                            sincospif((float)(ii + jj + k) * A.F1.inv_n, &sn, &cs);
                            a.x += vas[u].x * cs - vas[u].y * sn;
                            a.y += vas[u].x * sn + vas[u].y * cs;
                        }
                        a.x *= A.F1.scale;
                        a.y *= A.F1.scale;
                        float ww = 1.0f;
                        if (A.F1.W) {
                            ww = (Wi * A.F1.W[j]) * Wk;
                            a.x = __fdiv_rn(a.x, ww);
                            a.y = __fdiv_rn(a.y, ww);
                        }
                        if (A.f2) {
                            if (A.F2.fs) {
                                b.x += vbs[u].x * cs - vbs[u].y * sn;
                                b.y += vbs[u].x * sn + vbs[u].y * cs;
                            }
                            b.x *= A.F2.scale;
                            b.y *= A.F2.scale;
                            if (A.F2.W) {
                                b.x = __fdiv_rn(b.x, ww);
                                b.y = __fdiv_rn(b.y, ww);
                            }
                        }
                    }
                    val = (A.finish && use_tab) ? a.x : (A.f2 ? (a.x * b.x + a.y * b.y) : (a.x * a.x + a.y * a.y));
                }
                const float km2 = kmag2[u];
                const float mu2 = km2 > 0.0f ? __fdiv_rn(k2f, km2) : 0.0f;
                // incremental bin tracking: bk = #{b in 1..Nk : e[b] < km2}
                while (bk < Nk - 1 && km2 > s_ke[bk + 1]) bk++;
                while (bk > 0 && !(km2 > s_ke[bk])) bk--;
                while (bmu < Nmu - 1 && mu2 > s_me[bmu + 1]) bmu++;
                while (bmu > 0 && !(mu2 > s_me[bmu])) bmu--;
                const int key = bk * Nmu + bmu;
                if (key != acc.key) {
                    lane_flush<NPN>(A, acc, rep_off_bins, rep_off_poles, s_pidx);
                    acc.key = key; acc.bk = bk; acc.cnt = 0; acc.p = 0.0f; acc.k = 0.0f;
#pragma unroll
                    for (int q = 0; q < (NPN > 0 ? NPN : 1); q++) acc.pl[q] = 0.0f;
                }
                const float pv = mult * val;
                acc.cnt += cmult;
                acc.p += pv;
                acc.k = fmaf(mult, sqrtf(km2), acc.k);
                if (NPN > 0) {
                    const float sarg = A.even_only ? mu2 : sqrtf(mu2);
#pragma unroll
                    for (int q = 0; q < NPN; q++) {
                        if (q >= A.Npn) break;
                        const float *c = s_coef + q * ABK_POLE_NCOEF;
                        float pw;
                        if (A.even_only) {
                            pw = c[10];
                            pw = fmaf(pw, sarg, c[8]); pw = fmaf(pw, sarg, c[6]); pw = fmaf(pw, sarg, c[4]);
                            pw = fmaf(pw, sarg, c[2]); pw = fmaf(pw, sarg, c[0]);
                        } else {
                            pw = c[10];
#pragma unroll
                            for (int m = 9; m >= 0; m--) pw = fmaf(pw, sarg, c[m]);
                        }
                        acc.pl[q] = fmaf(pv, pw, acc.pl[q]);
                    }
                }
            }
        }
        lane_flush<NPN>(A, acc, rep_off_bins, rep_off_poles, s_pidx);
    }
}

// ------------------------------------------------------------------------------------------
// (k,mu) binning of a PENCIL (all x, a range of y, all k_z: what a rank of the sharded path holds after the transpose), fused
// finish step.  The two modes (+-i', j, k) share |k|^2, mu^2, the bin, the Legendre weights and -- for a symmetric window
// table -- the window product, and both live on this rank: a warp owns (|i'|, 32 k), walks the local j two rows at a time,
// loads both mirror entries of every mesh, applies the interlacing phase as a product of unit phasors (per-lane e^{i pi (+-i' + k)/n},
// per-row table e^{i pi j'/n}) and does the bin arithmetic ONCE per pair: half the divisions, square roots, bin searches,
// Legendre sums and reductions of power_bin2_kernel.  (The full-mesh kernel below shares four modes; +-j' sit on different
// ranks here.)
template <int NPN>
__global__ void __launch_bounds__(256) power_bin_pair_kernel(BinArgs A, unsigned *__restrict__ task_counter, int nrep)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_ke = reinterpret_cast<float *>(smem_raw);
    float *s_me = s_ke + (A.Nk + 1);
    float *s_coef = s_me + (A.Nmu + 1);
    int *s_pidx = reinterpret_cast<int *>(s_coef + A.Npn * ABK_POLE_NCOEF);
    const int lane = threadIdx.x & 31;
    const abk_kmesh M = A.M;
    const int n = M.n, nj = M.j1 - M.j0, amax = n - n / 2;
    const int hdr = (A.Nk + 1) + (A.Nmu + 1) + A.Npn * ABK_POLE_NCOEF + A.Npn;
    float2 *s_phj = reinterpret_cast<float2 *>(smem_raw + (size_t)((hdr + 3) & ~3) * 4);
    float *s_Wj = reinterpret_cast<float *>(s_phj + nj);

    for (int t = threadIdx.x; t <= A.Nk; t += blockDim.x) s_ke[t] = A.kedges2[t];
    for (int t = threadIdx.x; t <= A.Nmu; t += blockDim.x) s_me[t] = A.muedges2[t];
    for (int t = threadIdx.x; t < nj; t += blockDim.x) {
        float sn, cs;
        sincospif((float)fold(M.j0 + t, n) * A.F1.inv_n, &sn, &cs);
        s_phj[t] = make_float2(cs, sn);
        s_Wj[t] = A.F1.W ? A.F1.W[M.j0 + t] : 1.0f;
    }
    if (threadIdx.x == 0) {
        int q = 0;
        for (int p = 0; p < A.Np; p++)
            if (A.pole_ell[p] != 0) {
                for (int c = 0; c < ABK_POLE_NCOEF; c++) s_coef[q * ABK_POLE_NCOEF + c] = A.pole_coef[p * ABK_POLE_NCOEF + c];
                s_pidx[q] = p;
                q++;
            }
    }
    __syncthreads();

    const int Nk = A.Nk, Nmu = A.Nmu;
    const int nchunks = (M.nzc + 31) / 32;
    const unsigned ntasks = (unsigned)(amax + 1) * nchunks;
    const float e_lo = s_ke[0], e_hi = s_ke[Nk];
    const int rep = (nrep > 1) ? (int)(blockIdx.x % nrep) : 0;
    const size_t rep_off_bins = (size_t)rep * Nk * Nmu, rep_off_poles = (size_t)rep * A.Np * Nk;
    const bool inter1 = A.F1.fs != nullptr, cross = A.f2 != nullptr, inter2 = cross && A.F2.fs != nullptr;
    const float pscale = cross ? A.F1.scale * A.F2.scale : A.F1.scale * A.F1.scale;
    const int wpow = (A.F1.W ? 1 : 0) + ((cross ? A.F2.W : A.F1.W) ? 1 : 0);

    for (;;) {
        unsigned task = 0;
        if (lane == 0) task = atomicAdd(task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= ntasks) break;
        const int ai = task / nchunks, k0 = (task % nchunks) * 32;
        const int k = k0 + lane;
        const bool k_ok = k < M.nzc;
        const int ik2 = ai * ai + k * k;
        if ((float)(ai * ai + k0 * k0) >= e_hi) continue;  // the whole column pair lies beyond the last edge
        // members of |i'| = ai: i = ai (i' = +ai, if ai < n/2) and i = n - ai (i' = -ai, if ai >= 1); i' = -n/2 has no mirror
        const int nmem = (ai < n / 2 ? 1 : 0) + (ai >= 1 ? 1 : 0);
        const int i_a = (ai < n / 2) ? ai : n - ai, i_b = n - ai;
        const int64_t off_a = (int64_t)(i_a - M.i0) * M.stride_i + k, off_b = (int64_t)(i_b - M.i0) * M.stride_i + k;
        float2 e_a = make_float2(1.0f, 0.0f), e_b = make_float2(1.0f, 0.0f);  // e^{i pi (i' + k) / n} of the two members
        if (inter1 || inter2) {
            sincospif((float)(fold(i_a, n) + k) * A.F1.inv_n, &e_a.y, &e_a.x);
            sincospif((float)(fold(i_b, n) + k) * A.F1.inv_n, &e_b.y, &e_b.x);
        }
        const float k2f = (float)(k * k);
        const float Wi = A.F1.W ? A.F1.W[i_a] : 1.0f;      // == W[i_b]: the window table is symmetric
        const float Wk = (A.F1.W && k_ok) ? A.F1.W[k] : 1.0f;
        const float mult = (k == 0) ? 1.0f : 2.0f;
        const unsigned cmult = ((k == 0) ? 1u : 2u) * (unsigned)nmem;

        LaneAcc<NPN> acc;
        acc.key = -1; acc.bk = 0; acc.cnt = 0; acc.p = 0.0f; acc.k = 0.0f;
#pragma unroll
        for (int q = 0; q < (NPN > 0 ? NPN : 1); q++) acc.pl[q] = 0.0f;
        int bk = 0, bmu = 0;

        constexpr int ROWS = 2;
        for (int jl0 = 0; jl0 < nj; jl0 += ROWS) {
            // ---- issue the loads of ROWS rows x 2 members first -------------------------------------------
            float2 va[ROWS][2], vas[ROWS][2], vb[ROWS][2], vbs[ROWS][2];
            float kmag2[ROWS];
            bool use[ROWS];
#pragma unroll
            for (int u = 0; u < ROWS; u++) {
                const int jl = jl0 + u;
                const int jj = fold(M.j0 + jl, n);
                kmag2[u] = (float)(ik2 + jj * jj);
                use[u] = k_ok && jl < nj && kmag2[u] >= e_lo && kmag2[u] < e_hi;
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    va[u][mi] = vas[u][mi] = vb[u][mi] = vbs[u][mi] = make_float2(0.0f, 0.0f);
                    if (use[u] && mi < nmem) {
                        const int64_t idx = (mi == 0 ? off_a : off_b) + (int64_t)jl * M.stride_j;
                        va[u][mi] = __ldcs(A.f1 + idx);
                        if (inter1) vas[u][mi] = __ldcs(A.F1.fs + idx);
                        if (cross) {
                            vb[u][mi] = __ldcs(A.f2 + idx);
                            if (inter2) vbs[u][mi] = __ldcs(A.F2.fs + idx);
                        }
                    }
                }
            }
            // ---- per-row arithmetic, once for the pair ----------------------------------------------------------
#pragma unroll
            for (int u = 0; u < ROWS; u++) {
                if (!use[u]) continue;
                const int jl = jl0 + u;
                const float2 ej = s_phj[jl];
                float sum = 0.0f;
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    if (mi >= nmem) break;
                    const float2 e = mi == 0 ? e_a : e_b;
                    const float cs = e.x * ej.x - e.y * ej.y, sn = e.x * ej.y + e.y * ej.x;
                    float2 a = va[u][mi], b = vb[u][mi];
                    if (inter1) {
                        a.x += vas[u][mi].x * cs - vas[u][mi].y * sn;
                        a.y += vas[u][mi].x * sn + vas[u][mi].y * cs;
                    }
                    if (inter2) {
                        b.x += vbs[u][mi].x * cs - vbs[u][mi].y * sn;
                        b.y += vbs[u][mi].x * sn + vbs[u][mi].y * cs;
                    }
                    sum += cross ? (a.x * b.x + a.y * b.y) : (a.x * a.x + a.y * a.y);
                }
                float val = sum * pscale;
                if (wpow) {
                    const float ww = (Wi * s_Wj[jl]) * Wk;
                    val = __fdiv_rn(val, wpow == 2 ? ww * ww : ww);
                }
                const float km2 = kmag2[u];
                const float mu2 = km2 > 0.0f ? __fdiv_rn(k2f, km2) : 0.0f;
                while (bk < Nk - 1 && km2 > s_ke[bk + 1]) bk++;
                while (bk > 0 && !(km2 > s_ke[bk])) bk--;
                while (bmu < Nmu - 1 && mu2 > s_me[bmu + 1]) bmu++;
                while (bmu > 0 && !(mu2 > s_me[bmu])) bmu--;
                const int key = bk * Nmu + bmu;
                if (key != acc.key) {
                    lane_flush<NPN>(A, acc, rep_off_bins, rep_off_poles, s_pidx);
                    acc.key = key; acc.bk = bk; acc.cnt = 0; acc.p = 0.0f; acc.k = 0.0f;
#pragma unroll
                    for (int q = 0; q < (NPN > 0 ? NPN : 1); q++) acc.pl[q] = 0.0f;
                }
                const float pv = mult * val;
                acc.cnt += cmult;
                acc.p += pv;
                acc.k = fmaf(mult * (float)nmem, sqrtf(km2), acc.k);
                if (NPN > 0) {
                    const float sarg = A.even_only ? mu2 : sqrtf(mu2);
#pragma unroll
                    for (int q = 0; q < NPN; q++) {
                        if (q >= A.Npn) break;
                        const float *c = s_coef + q * ABK_POLE_NCOEF;
                        float pw;
                        if (A.even_only) {
                            pw = c[10];
                            pw = fmaf(pw, sarg, c[8]); pw = fmaf(pw, sarg, c[6]); pw = fmaf(pw, sarg, c[4]);
                            pw = fmaf(pw, sarg, c[2]); pw = fmaf(pw, sarg, c[0]);
                        } else {
                            pw = c[10];
#pragma unroll
                            for (int m = 9; m >= 0; m--) pw = fmaf(pw, sarg, c[m]);
                        }
                        acc.pl[q] = fmaf(pv, pw, acc.pl[q]);
                    }
                }
            }
        }
        lane_flush<NPN>(A, acc, rep_off_bins, rep_off_poles, s_pidx);
    }
}

// ------------------------------------------------------------------------------------------
// (k,mu) binning, symmetric formulation (single GPU, full mesh).
//
// The modes (+-i', +-j', k) share |k|^2, mu^2, hence the bin, the Legendre weights and the window
// product.  A warp owns (|i'|, 32 consecutive k) and walks |j'|: per step it loads the up to four
// mirror entries of each mesh, applies the interlacing phases as products of three unit phasors
// (table e^{i pi m / n} in shared memory: no sincos in the loop), and does the bin arithmetic ONCE
// for the group.  That is ~4x fewer bin computations and ~4x fewer reductions than one mode at a time.
// Requires W[n-a] == W[a] (checked on the host; true for the reference's tables).
template <int NPN>
__global__ void __launch_bounds__(256) power_bin_sym_kernel(BinArgs A, unsigned *__restrict__ task_counter, int nrep)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_ke = reinterpret_cast<float *>(smem_raw);
    float *s_me = s_ke + (A.Nk + 1);
    float *s_coef = s_me + (A.Nmu + 1);
    int *s_pidx = reinterpret_cast<int *>(s_coef + A.Npn * ABK_POLE_NCOEF);
    const int hdr = (A.Nk + 1) + (A.Nmu + 1) + A.Npn * ABK_POLE_NCOEF + A.Npn;
    const abk_kmesh M = A.M;
    const int n = M.n;
    const int amax = n - n / 2;  // largest |i'|
    float2 *s_ph = reinterpret_cast<float2 *>(smem_raw + (size_t)((hdr + 3) & ~3) * 4);  // e^{i pi m/n}, m = 0..amax
    float *s_W = reinterpret_cast<float *>(s_ph + (amax + 1));                          // W[m], m = 0..amax
    const int lane = threadIdx.x & 31;

    for (int t = threadIdx.x; t <= A.Nk; t += blockDim.x) s_ke[t] = A.kedges2[t];
    for (int t = threadIdx.x; t <= A.Nmu; t += blockDim.x) s_me[t] = A.muedges2[t];
    for (int t = threadIdx.x; t <= amax; t += blockDim.x) {
        float sn, cs;
        sincospif((float)t / (float)n, &sn, &cs);
        s_ph[t] = make_float2(cs, sn);
        // |i'| = t lives at index t (t < n/2) or n - t
        s_W[t] = (A.finish && A.F1.W) ? A.F1.W[t < n / 2 ? t : n - t] : 1.0f;
    }
    if (threadIdx.x == 0) {
        int q = 0;
        for (int p = 0; p < A.Np; p++)
            if (A.pole_ell[p] != 0) {
                for (int c = 0; c < ABK_POLE_NCOEF; c++) s_coef[q * ABK_POLE_NCOEF + c] = A.pole_coef[p * ABK_POLE_NCOEF + c];
                s_pidx[q] = p;
                q++;
            }
    }
    __syncthreads();

    const int Nk = A.Nk, Nmu = A.Nmu;
    const int nchunks = (M.nzc + 31) / 32;
    // A task is (|i'|, 32 k, a segment of BIN_AJ_SEG |j'|): a whole column would be up to n/2 + 1 dependent steps, and with
    // fewer columns than resident warps the kernel lasted as long as its longest column (513 steps of one memory latency
    // each at nmesh 1024).  Segments give every warp the same short walk; those beyond the last edge exit at once.
    const int nseg = amax / BIN_AJ_SEG + 1;
    const unsigned ntasks = (unsigned)(amax + 1) * nchunks * nseg;
    const float e_lo = s_ke[0], e_hi = s_ke[Nk];
    const int rep = (nrep > 1) ? (int)(blockIdx.x % nrep) : 0;
    const size_t rep_off_bins = (size_t)rep * Nk * Nmu, rep_off_poles = (size_t)rep * A.Np * Nk;
    const bool inter1 = A.finish && A.F1.fs != nullptr;
    const bool cross = A.f2 != nullptr;
    const bool inter2 = cross && A.finish && A.F2.fs != nullptr;
    const float scale2 = A.finish ? A.F1.scale * A.F1.scale : 1.0f;  // |v|^2 scales with scale^2

    for (;;) {
        unsigned task = 0;
        if (lane == 0) task = atomicAdd(task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= ntasks) break;
        // heavy columns (small |i'|) first: tasks are handed out in order of increasing |i'|
        const int seg = task % nseg, col = task / nseg;
        const int ai = col / nchunks, k0 = (col % nchunks) * 32;
        const int aj_lo = seg * BIN_AJ_SEG, aj_hi = min(amax, aj_lo + BIN_AJ_SEG - 1);
        const int k = k0 + lane;
        const bool k_ok = k < M.nzc;
        const int ik2 = ai * ai + k * k;
        if ((float)(ai * ai + k0 * k0 + aj_lo * aj_lo) >= e_hi) continue;
        const float k2f = (float)(k * k);
        const float mult = (k == 0) ? 1.0f : 2.0f;
        const unsigned cmult = (k == 0) ? 1u : 2u;
        // members of |i'| = ai: i = ai (if ai < n/2) and i = n - ai (if ai >= 1); same for j
        const int ni_mem = (ai < n / 2 ? 1 : 0) + ((ai >= 1 && ai <= amax) ? 1 : 0);
        const int i_first = (ai < n / 2) ? ai : n - ai;
        const int i_second = n - ai;  // used when ni_mem == 2
        // phasors: e^{i pi k/n} * e^{+- i pi ai/n}
        const float2 ek = k_ok ? s_ph[min(k, amax)] : make_float2(1.f, 0.f);
        const float2 ea = s_ph[ai];
        // sign of i' for the first member: + if ai < n/2 else -
        const float sgn1 = (ai < n / 2) ? 1.0f : -1.0f;
        const float2 eki1 = make_float2(ek.x * ea.x - sgn1 * ek.y * ea.y, ek.y * ea.x + sgn1 * ek.x * ea.y);
        const float2 eki2 = make_float2(ek.x * ea.x + ek.y * ea.y, ek.y * ea.x - ek.x * ea.y);  // i' = -ai
        const float Wik = s_W[ai] * 1.0f;
        const float Wk = s_W[min(k, amax)];

        LaneAcc<NPN> acc;
        acc.key = -1; acc.bk = 0; acc.cnt = 0; acc.p = 0.0f; acc.k = 0.0f;
#pragma unroll
        for (int q = 0; q < (NPN > 0 ? NPN : 1); q++) acc.pl[q] = 0.0f;
        int bk = 0, bmu = 0;

        // software prefetch of a later step's rows into the L1: the loop is otherwise latency-bound (one dependent batch of
        // loads per step)
        auto prefetch_step = [&](int an) {
            if (!k_ok || an > aj_hi) return;
            const int njn = (an < n / 2 ? 1 : 0) + ((an >= 1 && an <= amax) ? 1 : 0);
            const int jn1 = (an < n / 2) ? an : n - an, jn2 = n - an;
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                if (mi >= ni_mem) break;
                const int i = mi == 0 ? i_first : i_second;
#pragma unroll
                for (int mj = 0; mj < 2; mj++) {
                    if (mj >= njn) break;
                    const int64_t idx = (int64_t)i * M.stride_i + (int64_t)(mj == 0 ? jn1 : jn2) * M.stride_j + k;
                    if (A.real_in) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.real_in + idx));
                    } else {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(A.f1 + idx));
                        if (inter1) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.F1.fs + idx));
                        if (cross) {
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(A.f2 + idx));
                            if (inter2) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.F2.fs + idx));
                        }
                    }
                }
            }
        };
#pragma unroll
        for (int d = 0; d < BIN_PF_DIST; d++) prefetch_step(aj_lo + d);

        for (int aj = aj_lo; aj <= aj_hi; aj++) {
            const float km2 = (float)(ik2 + aj * aj);
            if ((float)(ai * ai + k0 * k0 + aj * aj) >= e_hi) break;  // |k| only grows with aj: warp-uniform exit
            const bool use = k_ok && km2 >= e_lo && km2 < e_hi;
            const int nj_mem = (aj < n / 2 ? 1 : 0) + ((aj >= 1 && aj <= amax) ? 1 : 0);
            const int j_first = (aj < n / 2) ? aj : n - aj;
            const int j_second = n - aj;
            const float sgj1 = (aj < n / 2) ? 1.0f : -1.0f;
            const float2 ej = s_ph[aj];
            float sum = 0.0f;
            int members = 0;
            prefetch_step(aj + BIN_PF_DIST);
            if (use) {
#pragma unroll
                for (int mi = 0; mi < 2; mi++) {
                    if (mi >= ni_mem) break;
                    const int i = mi == 0 ? i_first : i_second;
                    const float2 eki = mi == 0 ? eki1 : eki2;
#pragma unroll
                    for (int mj = 0; mj < 2; mj++) {
                        if (mj >= nj_mem) break;
                        const int j = mj == 0 ? j_first : j_second;
                        const float sg = mj == 0 ? sgj1 : -1.0f;
                        const int64_t idx = (int64_t)i * M.stride_i + (int64_t)j * M.stride_j + k;
                        float v;
                        if (A.real_in) {
                            v = __ldg(A.real_in + idx);
                        } else {
                            float2 a = __ldg(A.f1 + idx);
                            // phase = eki * e^{+- i pi aj / n}
                            const float2 ph = make_float2(eki.x * ej.x - sg * eki.y * ej.y, eki.y * ej.x + sg * eki.x * ej.y);
                            if (inter1) {
                                const float2 b = __ldg(A.F1.fs + idx);
                                a.x += b.x * ph.x - b.y * ph.y;
                                a.y += b.x * ph.y + b.y * ph.x;
                            }
                            if (cross) {
                                float2 c = __ldg(A.f2 + idx);
                                if (inter2) {
                                    const float2 d = __ldg(A.F2.fs + idx);
                                    c.x += d.x * ph.x - d.y * ph.y;
                                    c.y += d.x * ph.y + d.y * ph.x;
                                }
                                v = a.x * c.x + a.y * c.y;
                            } else {
                                v = a.x * a.x + a.y * a.y;
                            }
                        }
                        sum += v;
                        members++;
                    }
                }
            }
            if (!use) continue;
            float val = sum;
            if (A.finish) {
                val *= scale2;
                if (A.F1.W) {
                    const float ww = (Wik * s_W[aj]) * Wk;
                    val = __fdiv_rn(val, ww * ww);
                }
            }
            const float mu2 = km2 > 0.0f ? __fdiv_rn(k2f, km2) : 0.0f;
            while (bk < Nk - 1 && km2 > s_ke[bk + 1]) bk++;
            while (bk > 0 && !(km2 > s_ke[bk])) bk--;
            while (bmu < Nmu - 1 && mu2 > s_me[bmu + 1]) bmu++;
            while (bmu > 0 && !(mu2 > s_me[bmu])) bmu--;
            const int key = bk * Nmu + bmu;
            if (key != acc.key) {
                lane_flush<NPN>(A, acc, rep_off_bins, rep_off_poles, s_pidx);
                acc.key = key; acc.bk = bk; acc.cnt = 0; acc.p = 0.0f; acc.k = 0.0f;
#pragma unroll
                for (int q = 0; q < (NPN > 0 ? NPN : 1); q++) acc.pl[q] = 0.0f;
            }
            const float pv = mult * val;
            acc.cnt += cmult * (unsigned)members;
            acc.p += pv;
            acc.k = fmaf(mult * (float)members, sqrtf(km2), acc.k);
            if (NPN > 0) {
                const float sarg = A.even_only ? mu2 : sqrtf(mu2);
#pragma unroll
                for (int q = 0; q < NPN; q++) {
                    if (q >= A.Npn) break;
                    const float *c = s_coef + q * ABK_POLE_NCOEF;
                    float pw;
                    if (A.even_only) {
                        pw = c[10];
                        pw = fmaf(pw, sarg, c[8]); pw = fmaf(pw, sarg, c[6]); pw = fmaf(pw, sarg, c[4]);
                        pw = fmaf(pw, sarg, c[2]); pw = fmaf(pw, sarg, c[0]);
                    } else {
                        pw = c[10];
#pragma unroll
                        for (int m = 9; m >= 0; m--) pw = fmaf(pw, sarg, c[m]);
                    }
                    acc.pl[q] = fmaf(pv, pw, acc.pl[q]);
                }
            }
        }
        lane_flush<NPN>(A, acc, rep_off_bins, rep_off_poles, s_pidx);
    }
}

// fold the NREP replicas into the caller's sums (accumulating)
__global__ void __launch_bounds__(256) fold_replicas_kernel(const unsigned long long *__restrict__ rc,
                                                            const double *__restrict__ rp, const double *__restrict__ rk,
                                                            const double *__restrict__ rpl, int nrep, int64_t Nb,
                                                            int64_t Npl, unsigned long long *__restrict__ counts,
                                                            double *__restrict__ sum_p, double *__restrict__ sum_k,
                                                            double *__restrict__ sum_poles)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < Nb) {
        unsigned long long c = 0;
        double p = 0.0, k = 0.0;
        for (int r = 0; r < nrep; r++) { c += rc[r * Nb + t]; p += rp[r * Nb + t]; k += rk[r * Nb + t]; }
        counts[t] += c; sum_p[t] += p; sum_k[t] += k;
    }
    if (t < Npl) {
        double v = 0.0;
        for (int r = 0; r < nrep; r++) v += rpl[r * Npl + t];
        sum_poles[t] += v;
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

// Fused pack + transfer of the slab->pencil transpose: every (x, j) row of the local slab is written straight
// into its final place in the pencil buffer of the rank that owns j, through peer-mapped pointers (NVLink /
// NVSwitch P2P stores).  One read of the slab, no staging buffer, no separate collective.
struct PeerTable { float2 *p[MAX_RANKS]; };

__global__ void __launch_bounds__(256) transpose_scatter_p2p_kernel(const float2 *__restrict__ slab, PeerTable peers,
                                                                    int64_t nxl, int64_t ny, int64_t nzc, int nranks,
                                                                    SplitTable js, int64_t x_lo)
{
    __shared__ int64_t s_js[MAX_RANKS + 1];
    __shared__ float2 *s_peer[MAX_RANKS];
    for (int t = threadIdx.x; t <= nranks; t += blockDim.x) s_js[t] = js.v[t];
    for (int t = threadIdx.x; t < nranks; t += blockDim.x) s_peer[t] = peers.p[t];
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
        float2 *dst = s_peer[r] + ((x_lo + x) * nyl + (j - s_js[r])) * nzc;
        // 16-byte peer stores (two modes per lane and instruction): rows are nzc * 8 bytes long, so every other row starts
        // 8 bytes off a 16-byte boundary -- one leading mode goes alone there, and one trailing mode where the rest is odd
        const int64_t k0 = (reinterpret_cast<uintptr_t>(dst) & 8) ? 1 : 0;
        if (k0 && lane == 0) dst[0] = src[0];
        const int64_t npair = (nzc - k0) >> 1;
        for (int64_t p = lane; p < npair; p += 32) {
            const int64_t k = k0 + 2 * p;
            const float2 a = src[k], b = src[k + 1];
            *reinterpret_cast<float4 *>(dst + k) = make_float4(a.x, a.y, b.x, b.y);
        }
        if (((nzc - k0) & 1) && lane == 31) dst[nzc - 1] = src[nzc - 1];
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

// real (n,n,nzc) float32 -> complex64 with zero imaginary part (input of the C2R transform in pk_to_xi)
__global__ void __launch_bounds__(256) real_to_complex_kernel(const float *__restrict__ in, float2 *__restrict__ out,
                                                              int64_t size)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < size; t += (int64_t)gridDim.x * blockDim.x)
        out[t] = make_float2(in[t], 0.0f);
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
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
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

static int finish_impl(abk_ctx *ctx, const abk_kmesh *mesh_h, void *f, const void *fs, const float *W, float scale, float phase_step)
{
    int rc = check_mesh(*mesh_h);
    if (rc) return rc;
    FinishArgs F;
    F.fs = (const float2 *)fs;
    F.W = W;
    F.scale = scale;
    F.n = mesh_h->n;
    F.inv_n = phase_step;  // phase of mode (i', j', k) = pi * (i' + j' + k) * phase_step
    const int64_t nrows = (int64_t)(mesh_h->i1 - mesh_h->i0) * (mesh_h->j1 - mesh_h->j0);
    if (nrows == 0) return ABK_OK;
    ABK_LAUNCH(ctx, ABK_K_FINISH,
               finish_kernel<<<grid_for(ctx, nrows * 32, 256, 8), 256, 0, ctx->stream>>>((float2 *)f, F, *mesh_h));
    return ABK_OK;
}

extern "C" int abk_field_fft_finish(abk_ctx *ctx, const abk_kmesh *mesh_h, void *f, const void *fs, const float *W,
                                    float scale)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && mesh_h && f, "abk_field_fft_finish: null argument");
    return finish_impl(ctx, mesh_h, f, fs, W, scale, 1.0f / (float)mesh_h->n);
}

extern "C" int abk_shift_field_fft(abk_ctx *ctx, const abk_kmesh *mesh_h, void *f, const void *fs, double d_over_L, float scale)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && mesh_h && f && fs, "abk_shift_field_fft: null argument");
    return finish_impl(ctx, mesh_h, f, fs, nullptr, scale, (float)d_over_L);
}

extern "C" int abk_raw_power(abk_ctx *ctx, const void *f1, const void *f2, float *out, int64_t size)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && f1 && out && size >= 0, "abk_raw_power: bad arguments");
    if (size == 0) return ABK_OK;
    ABK_LAUNCH(ctx, ABK_K_RAW_POWER, raw_power_kernel<<<grid_for(ctx, size, 256, 16), 256, 0, ctx->stream>>>((const float2 *)f1, (const float2 *)f2, out, size));
    return ABK_OK;
}

extern "C" int abk_real_to_complex(abk_ctx *ctx, const float *in, void *out, int64_t size)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && in && out && size >= 0, "abk_real_to_complex: bad arguments");
    if (size == 0) return ABK_OK;
    ABK_LAUNCH(ctx, ABK_K_MISC, real_to_complex_kernel<<<grid_for(ctx, size, 256, 16), 256, 0, ctx->stream>>>(in, (float2 *)out, size));
    return ABK_OK;
}

extern "C" int abk_power_bin(abk_ctx *ctx, const abk_bin_request *R)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
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
    const int64_t Nb = (int64_t)A.Nk * A.Nmu, Npl = (int64_t)A.Np * A.Nk;
    const int hdr = (A.Nk + 1) + (A.Nmu + 1) + A.Npn * ABK_POLE_NCOEF + A.Npn;
    const size_t hdr_bytes = (size_t)((hdr + 3) & ~3) * 4;
    ABK_REQUIRE(hdr_bytes + 1024 < (size_t)ctx->smem_optin, "abk_power_bin: edge tables (%zu B) do not fit in shared memory",
                hdr_bytes);

    // replicated sums in the caller's scratch (optional): [counts | sum_p | sum_k | sum_poles] x nrep
    int nrep = 1;
    const size_t rep_bytes = (size_t)(3 * Nb + Npl) * 8;
    if (R->scratch && R->scratch_bytes >= rep_bytes * BIN_NREP) nrep = BIN_NREP;
    unsigned long long *caller_counts = A.counts;
    double *caller_p = A.sum_p, *caller_k = A.sum_k, *caller_pl = A.sum_poles;
    if (nrep > 1) {
        ABK_CHECK_CUDA(cudaMemsetAsync(R->scratch, 0, rep_bytes * nrep, ctx->stream));
        char *base = (char *)R->scratch;
        A.counts = (unsigned long long *)base;
        A.sum_p = (double *)(base + (size_t)nrep * Nb * 8);
        A.sum_k = (double *)(base + (size_t)2 * nrep * Nb * 8);
        A.sum_poles = (double *)(base + (size_t)3 * nrep * Nb * 8);
    }
    unsigned *task_counter = (unsigned *)(ctx->d_scalars + 2);
    ABK_CHECK_CUDA(cudaMemsetAsync(task_counter, 0, sizeof(unsigned), ctx->stream));
    const int blocks = ctx->num_sms * 4;
    void (*kern)(BinArgs, unsigned *, int) = nullptr;
    void (*kern2)(BinArgs, unsigned *, int, int) = nullptr;
    int use_tab = 0;
    // mirror-symmetric kernel: needs the full mesh on this GPU, the finish step fused (so the scale and
    // window can be applied to the group sum) or no finish at all, and a symmetric window table
    const int n = A.M.n, amax = n - n / 2;
    const bool full = A.M.i0 == 0 && A.M.i1 == n && A.M.j0 == 0 && A.M.j1 == n && A.M.nzc == n / 2 + 1;
    const bool same_scale = !A.f2 || (A.F1.scale == A.F2.scale && A.F1.W == A.F2.W);
    bool sym = full && same_scale && R->w_symmetric && n >= 4 && !ctx->bin_no_sym;
    size_t smem = hdr_bytes;
    if (sym) {
        smem = hdr_bytes + (size_t)(amax + 1) * 12 + 16;
        if (smem + 1024 > (size_t)ctx->smem_optin) sym = false;
    }
    if (sym) {
        if (A.Npn == 0) kern = power_bin_sym_kernel<0>;
        else if (A.Npn <= 2) kern = power_bin_sym_kernel<2>;
        else if (A.Npn <= 4) kern = power_bin_sym_kernel<4>;
        else kern = power_bin_sym_kernel<ABK_MAX_POLES>;
    } else {
        smem = hdr_bytes;
        // fused finish step: per-row phasor and window tables in shared memory if they fit (else sincos per mode)
        const size_t tab = (size_t)(A.M.j1 - A.M.j0) * 12 + 16;
        if (A.finish && hdr_bytes + tab + 1024 <= (size_t)ctx->smem_optin) {
            use_tab = 1;
            smem = hdr_bytes + tab;
        }
        // pencil of a sharded mesh (all x-planes here): the +-i' mirror pair shares the bin arithmetic
        const bool all_x = A.M.i0 == 0 && A.M.i1 == n;
        if (use_tab && all_x && same_scale && R->w_symmetric && n >= 4 && !ctx->bin_no_sym) {
            if (A.Npn == 0) kern = power_bin_pair_kernel<0>;
            else if (A.Npn <= 2) kern = power_bin_pair_kernel<2>;
            else if (A.Npn <= 4) kern = power_bin_pair_kernel<4>;
            else kern = power_bin_pair_kernel<ABK_MAX_POLES>;
        } else if (A.Npn == 0) kern2 = power_bin2_kernel<0>;
        else if (A.Npn <= 2) kern2 = power_bin2_kernel<2>;
        else if (A.Npn <= 4) kern2 = power_bin2_kernel<4>;
        else kern2 = power_bin2_kernel<ABK_MAX_POLES>;
    }
    if (kern2) {
        ABK_CHECK_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ABK_LAUNCH(ctx, ABK_K_POWER_BIN, kern2<<<blocks, 256, smem, ctx->stream>>>(A, task_counter, nrep, use_tab));
    } else {
        ABK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ABK_LAUNCH(ctx, ABK_K_POWER_BIN, kern<<<blocks, 256, smem, ctx->stream>>>(A, task_counter, nrep));
    }
    if (nrep > 1) {
        const int64_t m = Nb > Npl ? Nb : Npl;
        ABK_LAUNCH(ctx, ABK_K_MISC,
                   fold_replicas_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(
                       A.counts, A.sum_p, A.sum_k, A.sum_poles, nrep, Nb, Npl, caller_counts, caller_p, caller_k, caller_pl));
    }
    return ABK_OK;
}

extern "C" int abk_power_bin_scratch_bytes(int Nk, int Nmu, int Np, size_t *bytes)
{
    ABK_REQUIRE(bytes && Nk >= 1 && Nmu >= 1 && Np >= 0, "abk_power_bin_scratch_bytes: bad arguments");
    *bytes = (size_t)(3 * (int64_t)Nk * Nmu + (int64_t)Np * Nk) * 8 * BIN_NREP;
    return ABK_OK;
}

extern "C" int abk_add_planes(abk_ctx *ctx, float *dst, const float *src, int64_t nplanes, int64_t ny, int64_t nz,
                              int64_t ldz)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && dst && src && nplanes >= 0 && ny > 0 && nz > 0 && ldz >= nz, "abk_add_planes: bad arguments");
    if (nplanes == 0) return ABK_OK;
    ABK_LAUNCH(ctx, ABK_K_ADD_PLANES,
               add_planes_kernel<<<grid_for(ctx, nplanes * ny * ldz, 256, 16), 256, 0, ctx->stream>>>(dst, src, nplanes * ny, nz, ldz));
    return ABK_OK;
}

extern "C" int abk_transpose_scatter_p2p(abk_ctx *ctx, const void *slab, void *const *peer_pencils_h, int64_t nxl,
                                         int64_t ny, int64_t nzc, int nranks, const int64_t *jsplit_h, int64_t x_lo)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && slab && peer_pencils_h && nxl >= 0 && ny > 0 && nzc > 0 && nranks > 0 && nranks <= MAX_RANKS && jsplit_h,
                "abk_transpose_scatter_p2p: bad arguments");
    ABK_REQUIRE(jsplit_h[0] == 0 && jsplit_h[nranks] == ny, "abk_transpose_scatter_p2p: jsplit must run from 0 to ny");
    if (nxl == 0) return ABK_OK;
    SplitTable js;
    PeerTable pt;
    for (int r = 0; r <= nranks; r++) js.v[r] = jsplit_h[r];
    for (int r = 0; r < nranks; r++) {
        ABK_REQUIRE(peer_pencils_h[r] != nullptr, "abk_transpose_scatter_p2p: null peer pointer for rank %d", r);
        pt.p[r] = (float2 *)peer_pencils_h[r];
    }
    ABK_LAUNCH(ctx, ABK_K_TRANSPOSE_PACK, transpose_scatter_p2p_kernel<<<grid_for(ctx, nxl * ny * 32, 256, 8), 256, 0, ctx->stream>>>(
                                              (const float2 *)slab, pt, nxl, ny, nzc, nranks, js, x_lo));
    return ABK_OK;
}

extern "C" int abk_transpose_pack(abk_ctx *ctx, const void *slab, void *sendbuf, int64_t nxl, int64_t ny, int64_t nzc,
                                  int nranks, const int64_t *jsplit_h)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
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
