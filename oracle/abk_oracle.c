/*
 * abk_oracle.c -- CPU restatement of the abacusutils TSC + power-spectrum path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path
 * in abacusutils_b200/csrc.  Nothing in the product package may import, link or
 * call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.  It restates, in plain C + OpenMP, the algorithm of
 * the reference's Numba kernels; each function cites the reference lines it
 * follows (paths relative to /root/reference).
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this restatement against
 *   - the analytic single-particle KAT of tests/test_tsc.py:25-90,
 *   - the reference's own golden grids tests/ref_tsc/{tsc,nbodykit_tsc}_ngrid{10,256}.asdf
 *     (decoded by tests/golden/make_golden.py into tests/golden/),
 *   - outputs of the unmodified reference (imported from /root/reference in the
 *     build container by tests/golden/make_golden.py) for bin_kmu / calc_power.
 *
 * Arithmetic is float32 wherever the reference's is (fields, weights, stencil
 * weights, kmag2/mu2, bin sums), so integer results are bit-identical and float
 * results differ only by summation order / FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float re, im; } c64;

static int clamp_threads(int nthread)
{
#ifdef _OPENMP
    if (nthread <= 0) nthread = omp_get_max_threads();
    return nthread;
#else
    (void)nthread;
    return 1;
#endif
}

int abko_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* T2: one-shot periodic wrap, in place.  analysis/tsc.py:219-226            */
/* The comparison is done against the double `box`; the store is float32,    */
/* so a value may land exactly on `box` (handled by the index wrap in T6).   */
void abko_wrap_inplace_f32(float *pos, int64_t n, double box, int nthread)
{
    nthread = clamp_threads(nthread);
    #pragma omp parallel for num_threads(nthread) schedule(static)
    for (int64_t i = 0; i < 3 * n; i++) {
        float v = pos[i];
        if ((double)v >= box) pos[i] = (float)((double)v - box);
        else if (v < 0.0f)    pos[i] = (float)((double)v + box);
    }
}

/* ------------------------------------------------------------------------- */
/* T3: counting sort of particles into `npart` stripes along `coord`.        */
/* analysis/tsc.py:259-384 (key :329-336; stable thread-major order :338-376) */
/* key = min(int32(pos[coord] * f32(npart/box)), npart-1).  Output order is   */
/* the stable partition.  keys_out may be NULL.                */
int abko_partition_f32(const float *pos, const float *w, int64_t N, int npart,
                       double box, int coord, float *psort, float *wsort,
                       int64_t *starts, int nthread)
{
    nthread = clamp_threads(nthread);
    const float inv_pwidth = (float)((double)npart / box);
    int32_t *keys = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    int64_t *counts = (int64_t *)calloc((size_t)nthread * npart, sizeof(int64_t));
    if (!keys || !counts) { free(keys); free(counts); return -1; }

    #pragma omp parallel num_threads(nthread)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
        int nt = omp_get_num_threads();
#else
        int t = 0, nt = 1;
#endif
        int64_t lo = N * t / nt, hi = N * (t + 1) / nt;
        int64_t *c = counts + (size_t)t * npart;
        for (int64_t i = lo; i < hi; i++) {
            int32_t k = (int32_t)(pos[3 * i + coord] * inv_pwidth);
            if (k > npart - 1) k = npart - 1;
            keys[i] = k;
            c[k]++;
        }
        #pragma omp barrier
        #pragma omp single
        {
            /* partition-major, thread-minor exclusive scan */
            int64_t run = 0;
            for (int p = 0; p < npart; p++) {
                starts[p] = run;
                for (int tt = 0; tt < nt; tt++) {
                    int64_t v = counts[(size_t)tt * npart + p];
                    counts[(size_t)tt * npart + p] = run;
                    run += v;
                }
            }
            starts[npart] = N;
        }
        for (int64_t i = lo; i < hi; i++) {
            int64_t s = c[keys[i]]++;
            psort[3 * s + 0] = pos[3 * i + 0];
            psort[3 * s + 1] = pos[3 * i + 1];
            psort[3 * s + 2] = pos[3 * i + 2];
            if (w) wsort[s] = w[i];
        }
    }
    free(keys);
    free(counts);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* T5/T6: serial 27-point TSC scatter.  analysis/tsc.py:387-391, 394-507      */
/* p = (pos + offset) * f32(g/box); i = rint(p) (half-to-even); d = i - p;    */
/* w0 = 0.75 - d^2, w- = 0.5(0.5+d)^2, w+ = 0.5(0.5-d)^2;                     */
/* index wrap: >= g -> -g (T5); negative -> +g (NumPy negative indexing).     */
static inline int wrap_idx(int i, int g)
{
    if (i >= g) i -= g;
    if (i < 0) i += g;
    return i;
}

static void tsc_scatter_range(const float *pos, const float *w, int64_t lo, int64_t hi,
                              float *dens, int gx, int gy, int gz, float inv_hx,
                              float inv_hy, float inv_hz, float off)
{
    const size_t sy = (size_t)gz, sx = (size_t)gy * gz;
    for (int64_t n = lo; n < hi; n++) {
        const float W = w ? w[n] : 1.0f;
        const float px = (pos[3 * n + 0] + off) * inv_hx;
        const float py = (pos[3 * n + 1] + off) * inv_hy;
        const float pz = (pos[3 * n + 2] + off) * inv_hz;
        const int ix = (int)rintf(px), iy = (int)rintf(py), iz = (int)rintf(pz);
        const float dx = (float)ix - px, dy = (float)iy - py, dz = (float)iz - pz;
        float wx[3], wy[3], wz[3];
        wx[0] = 0.5f * (0.5f + dx) * (0.5f + dx); wx[1] = 0.75f - dx * dx; wx[2] = 0.5f * (0.5f - dx) * (0.5f - dx);
        wy[0] = 0.5f * (0.5f + dy) * (0.5f + dy); wy[1] = 0.75f - dy * dy; wy[2] = 0.5f * (0.5f - dy) * (0.5f - dy);
        wz[0] = 0.5f * (0.5f + dz) * (0.5f + dz); wz[1] = 0.75f - dz * dz; wz[2] = 0.5f * (0.5f - dz) * (0.5f - dz);
        int jx[3], jy[3], jz[3];
        for (int a = 0; a < 3; a++) {
            jx[a] = wrap_idx(ix + a - 1, gx);
            jy[a] = wrap_idx(iy + a - 1, gy);
            jz[a] = wrap_idx(iz + a - 1, gz);
        }
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                float *row = dens + jx[a] * sx + jy[b] * sy;
                const float wab = wx[a] * wy[b];
                for (int c = 0; c < 3; c++) row[jz[c]] += wab * wz[c] * W;
            }
    }
}

void abko_tsc_scatter_f32(const float *pos, const float *w, int64_t N, float *dens,
                          int gx, int gy, int gz, double box, double offset)
{
    tsc_scatter_range(pos, w, 0, N, dens, gx, gy, gz, (float)(gx / box), (float)(gy / box),
                      (float)(gz / box), (float)offset);
}

/* CIC: analysis/cic.py:13-125 `cic_serial` (serial in the reference).  Same 27-cell update as TSC with  */
/* weights (max(d,0), 1-|d|, max(-d,0)); p = (pos / box) * g and all products evaluated in double, the  */
/* sum is stored into the float32 grid.                                                                */
void abko_cic_serial_f32(const float *pos, const float *w, int64_t N, float *dens, int gx, int gy, int gz,
                         double box)
{
    const size_t sy = (size_t)gz, sx = (size_t)gy * gz;
    for (int64_t n = 0; n < N; n++) {
        const double W = w ? (double)w[n] : 1.0;
        const double p[3] = {((double)pos[3 * n] / box) * gx, ((double)pos[3 * n + 1] / box) * gy,
                             ((double)pos[3 * n + 2] / box) * gz};
        const int g[3] = {gx, gy, gz};
        double wt[3][3];
        int idx[3][3];
        for (int a = 0; a < 3; a++) {
            const int i = (int)rint(p[a]);
            const double d = (double)i - p[a];
            wt[a][1] = 1.0 - fabs(d);
            if (d > 0.0) { wt[a][0] = d; wt[a][2] = 0.0; } else { wt[a][2] = -d; wt[a][0] = 0.0; }
            for (int q = 0; q < 3; q++) idx[a][q] = wrap_idx(i + q - 1, g[a]);
        }
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                float *row = dens + idx[0][a] * sx + idx[1][b] * sy;
                for (int c = 0; c < 3; c++) row[idx[2][c]] += (float)(wt[0][a] * wt[1][b] * wt[2][c] * W);
            }
    }
}

/* T4: two-colour stripe deposit.  analysis/tsc.py:229-256                    */
/* Even stripes in parallel, then odd stripes.                               */
void abko_tsc_stripes_f32(const float *psort, const float *wsort, const int64_t *starts,
                          int npart, float *dens, int gx, int gy, int gz, double box,
                          double offset, int nthread)
{
    nthread = clamp_threads(nthread);
    const float ihx = (float)(gx / box), ihy = (float)(gy / box), ihz = (float)(gz / box);
    const float off = (float)offset;
    for (int colour = 0; colour < 2; colour++) {
        #pragma omp parallel for num_threads(nthread) schedule(dynamic, 1)
        for (int s = colour; s < npart; s += 2)
            tsc_scatter_range(psort, wsort, starts[s], starts[s + 1], dens, gx, gy, gz, ihx, ihy,
                              ihz, off);
    }
}

/* ------------------------------------------------------------------------- */
/* P5: overdensity normalisation in place.  power_spectrum.py:860-901         */
void abko_normalize_field_f32(float *field, int64_t size, double tot_weight, int nthread)
{
    nthread = clamp_threads(nthread);
    const float norm = (float)((double)size / tot_weight);
    #pragma omp parallel for num_threads(nthread) schedule(static)
    for (int64_t i = 0; i < size; i++) field[i] = field[i] * norm - 1.0f;
}

/* P8 tail: flat scale of a complex array.  power_spectrum.py:1073-1078       */
void abko_scale_c64(c64 *f, int64_t size, float a, int nthread)
{
    nthread = clamp_threads(nthread);
    #pragma omp parallel for num_threads(nthread) schedule(static)
    for (int64_t i = 0; i < size; i++) { f[i].re *= a; f[i].im *= a; }
}

/* P7: interlacing combination, in place.  power_spectrum.py:904-948          */
/* f <- (f + fs * exp(i * f32(0.5 d) * (kx+ky+kz))) * f32(0.5/n^3); the phase */
/* argument is float32, the exponential is evaluated in double (complex128).  */
void abko_shift_field_fft(c64 *f, const c64 *fs, int n, double L, double d, int nthread)
{
    nthread = clamp_threads(nthread);
    const int kzlen = n / 2 + 1;
    const float dk = (float)(2.0 * M_PI / L);
    const float norm = (float)(0.5 / ((double)n * n * n));
    const double fac = (double)(float)(0.5 * (double)(float)d);
    #pragma omp parallel for num_threads(nthread) schedule(static)
    for (int i = 0; i < n; i++) {
        const float kx = (i < n / 2) ? (float)i * dk : (float)(i - n) * dk;
        for (int j = 0; j < n; j++) {
            const float ky = (j < n / 2) ? (float)j * dk : (float)(j - n) * dk;
            c64 *row = f + ((size_t)i * n + j) * kzlen;
            const c64 *rs = fs + ((size_t)i * n + j) * kzlen;
            for (int k = 0; k < kzlen; k++) {
                const float kz = (float)k * dk;
                const double arg = fac * (double)(kx + ky + kz);
                const double c = cos(arg), s = sin(arg);
                const double re = (double)row[k].re + ((double)rs[k].re * c - (double)rs[k].im * s);
                const double im = (double)row[k].im + ((double)rs[k].re * s + (double)rs[k].im * c);
                row[k].re = (float)re * norm;
                row[k].im = (float)im * norm;
            }
        }
    }
}

/* P8: window compensation, f /= (W_i*W_j)*W_k.  power_spectrum.py:1062-1070  */
void abko_compensate(c64 *f, const float *W, int n, int nthread)
{
    nthread = clamp_threads(nthread);
    const int kzlen = n / 2 + 1;
    #pragma omp parallel for num_threads(nthread) schedule(static)
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const float wij = W[i] * W[j];
            c64 *row = f + ((size_t)i * n + j) * kzlen;
            for (int k = 0; k < kzlen; k++) {
                const float ww = wij * W[k];
                row[k].re /= ww;
                row[k].im /= ww;
            }
        }
}

/* P11: |f|^2 or Re(conj(f1) f2).  power_spectrum.py:707-727                  */
void abko_raw_power(const c64 *f1, const c64 *f2, float *out, int64_t size, int nthread)
{
    nthread = clamp_threads(nthread);
    if (f2) {
        #pragma omp parallel for num_threads(nthread) schedule(static)
        for (int64_t i = 0; i < size; i++) out[i] = f1[i].re * f2[i].re + f1[i].im * f2[i].im;
    } else {
        #pragma omp parallel for num_threads(nthread) schedule(static)
        for (int64_t i = 0; i < size; i++) {
            /* np.abs(z)**2: hypot then square */
            const float a = hypotf(f1[i].re, f1[i].im);
            out[i] = a * a;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* P10: Legendre polynomial as a polynomial in x = mu^2.                      */
/* power_spectrum.py:29-147 (P_n :121-147)                                    */
static double factorial_d(int n)
{
    double r = 1.0;
    for (int i = 2; i <= n; i++) r *= i;
    return r;
}
static double n_choose_k_d(int n, int k)
{
    return floor(factorial_d(n) / (factorial_d(k) * factorial_d(n - k)) + 0.5);
}
float abko_P_n(float x, int n)
{
    float sum = 0.0f;
    for (int k = 0; k <= n / 2; k++) {
        const float factor = (float)(n_choose_k_d(n, k) * n_choose_k_d(2 * n - 2 * k, n));
        const float term = factor * powf(x, (float)(0.5 * (n - 2 * k)));
        if (k % 2 == 0) sum += term; else sum -= term;
    }
    sum *= (float)pow(0.5, n);
    return sum;
}

/* ------------------------------------------------------------------------- */
/* P9: (k,mu) binning with Legendre multipoles.  power_spectrum.py:150-300    */
/*                                                                           */
/*  kmag2 = f32(i'^2 + j'^2 + k^2);  mu2 = f32(k^2) / kmag2 (one rounding;    */
/*  0 at DC);  skip kmag2 < kedges2[0];  stop row at kmag2 >= kedges2[-1];    */
/*  bins are (lo, hi] found by monotone advance;  multiplicity 1 on the k=0   */
/*  plane, 2 elsewhere (Nyquist included).  Per-thread accumulators, summed   */
/*  at the end.  acc64 != 0 accumulates the float sums in double (a tighter   */
/*  checker for the GPU path); acc64 == 0 reproduces the reference's float32  */
/*  accumulation.  `weights` has row length `zlen` (n/2+1 for rfft layout,    */
/*  n for a real-space mesh) and only k < n/2+1 is visited.                   */
/*  Outputs are the RAW sums; the caller divides by the counts.               */
int abko_bin_kmu(int n, double L, const double *kedges, int Nk, const double *muedges, int Nmu,
                 const float *weights, int64_t zlen, const int64_t *poles, int Np, int fourier,
                 int acc64, int nthread, int64_t *counts /*Nk*Nmu*/, double *sum_w /*Nk*Nmu*/,
                 double *sum_k /*Nk*Nmu*/, double *sum_poles /*Np*Nk*/)
{
    nthread = clamp_threads(nthread);
    const int kzlen = n / 2 + 1;
    const double dk = fourier ? 2.0 * M_PI / L : L / n;
    float *kedges2 = (float *)malloc(sizeof(float) * (Nk + 1));
    float *muedges2 = (float *)malloc(sizeof(float) * (Nmu + 1));
    if (!kedges2 || !muedges2) return -1;
    for (int b = 0; b <= Nk; b++) { double t = kedges[b] / dk; kedges2[b] = (float)(t * t); }
    for (int b = 0; b <= Nmu; b++) muedges2[b] = (float)(muedges[b] * muedges[b]);

    const size_t nb = (size_t)Nk * Nmu, npb = (size_t)Np * Nk;
    memset(counts, 0, sizeof(int64_t) * nb);
    memset(sum_w, 0, sizeof(double) * nb);
    memset(sum_k, 0, sizeof(double) * nb);
    if (npb) memset(sum_poles, 0, sizeof(double) * npb);

    #pragma omp parallel num_threads(nthread)
    {
        int64_t *c_t = (int64_t *)calloc(nb ? nb : 1, sizeof(int64_t));
        double *w_d = (double *)calloc(nb ? nb : 1, sizeof(double));
        double *k_d = (double *)calloc(nb ? nb : 1, sizeof(double));
        double *p_d = (double *)calloc(npb ? npb : 1, sizeof(double));
        float *w_f = (float *)calloc(nb ? nb : 1, sizeof(float));
        float *k_f = (float *)calloc(nb ? nb : 1, sizeof(float));
        float *p_f = (float *)calloc(npb ? npb : 1, sizeof(float));

        #pragma omp for schedule(static)
        for (int i = 0; i < n; i++) {
            const int64_t ii = (i < n / 2) ? i : i - n;
            const int64_t i2 = ii * ii;
            for (int j = 0; j < n; j++) {
                const int64_t jj = (j < n / 2) ? j : j - n;
                const int64_t j2 = jj * jj;
                int bk = 0, bmu = 0;
                const float *row = weights + ((size_t)i * n + j) * (size_t)zlen;
                for (int k = 0; k < kzlen; k++) {
                    const float kmag2 = (float)(i2 + j2 + (int64_t)k * k);
                    const float mu2 = kmag2 > 0.0f ? (float)((int64_t)k * k) / kmag2 : 0.0f;
                    if (kmag2 < kedges2[0]) continue;
                    if (kmag2 >= kedges2[Nk]) break;
                    while (kmag2 > kedges2[bk + 1]) bk++;
                    while (bmu + 1 < Nmu && mu2 > muedges2[bmu + 1]) bmu++;
                    const size_t b = (size_t)bk * Nmu + bmu;
                    const float mult = (k == 0) ? 1.0f : 2.0f;
                    const float wv = row[k];
                    c_t[b] += (k == 0) ? 1 : 2;
                    if (acc64) {
                        w_d[b] += (double)mult * wv;
                        k_d[b] += (double)mult * (double)sqrtf(kmag2);
                    } else {
                        w_f[b] += mult * wv;
                        k_f[b] += (float)((double)mult * (double)sqrtf(kmag2) * dk);
                    }
                    for (int ip = 0; ip < Np; ip++) {
                        const int pole = (int)poles[ip];
                        if (pole == 0) continue;
                        const float pw = (float)(2 * pole + 1) * abko_P_n(mu2, pole);
                        if (acc64) p_d[(size_t)ip * Nk + bk] += (double)(mult * wv * pw);
                        else       p_f[(size_t)ip * Nk + bk] += mult * wv * pw;
                    }
                }
            }
        }
        #pragma omp critical
        {
            for (size_t b = 0; b < nb; b++) {
                counts[b] += c_t[b];
                sum_w[b] += acc64 ? w_d[b] : (double)w_f[b];
                sum_k[b] += acc64 ? k_d[b] * dk : (double)k_f[b];
            }
            for (size_t b = 0; b < npb; b++) sum_poles[b] += acc64 ? p_d[b] : (double)p_f[b];
        }
        free(c_t); free(w_d); free(k_d); free(p_d); free(w_f); free(k_f); free(p_f);
    }
    free(kedges2);
    free(muedges2);
    return 0;
}
