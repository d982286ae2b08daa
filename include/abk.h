/*
 * abk.h -- C ABI of libabk.so: the B200 (sm_100a) implementation of the abacusutils
 * density-field power-spectrum hot path.
 *
 * The reference (abacusorg/abacusutils) has no FFI on this path: its boundary is the Python
 * module API of abacusnbody.analysis.tsc / abacusnbody.analysis.power_spectrum, whose hot loops
 * are Numba @njit kernels.  Each entry point below replaces one of those kernels; the comment
 * above it cites the reference function it stands in for (paths relative to
 * /root/reference/abacusnbody/analysis/).  The Python shims in abacusutils_b200/analysis/ keep the
 * reference signatures and call these through ctypes (see INTEGRATION.md for the binding a
 * maintainer would add to the reference itself).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _h;
 *   - plain pointers and sizes only, no framework types;
 *   - every function returns 0 (ABK_OK) or a negative error code; abk_last_error() returns the
 *     message of the last failure on the calling thread;
 *   - work is enqueued on the context's stream (abk_ctx_set_stream) and is asynchronous unless
 *     stated otherwise; the library never allocates large device buffers behind the caller's
 *     back: scratch is sized by the *_scratch_bytes queries and passed in;
 *   - grids are C-ordered (x slowest, z contiguous) float32 with a z row length of `ldz` floats
 *     (ldz == nz for a plain grid, ldz == 2*(nz/2+1) for the in-place R2C layout).
 */
#ifndef ABK_H
#define ABK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABK_VERSION 100

#define ABK_OK 0
#define ABK_ERR_INVALID (-1) /* bad argument */
#define ABK_ERR_CUDA (-2)    /* CUDA runtime failure */
#define ABK_ERR_CUFFT (-3)   /* cuFFT failure */
#define ABK_ERR_SCRATCH (-4) /* scratch buffer too small */

#define ABK_MAX_SEGMENTS 16
#define ABK_MAX_POLES 16
#define ABK_POLE_NCOEF 11 /* polynomial in mu = sqrt(mu2) up to degree 10 (ell <= 10) */

typedef struct abk_ctx abk_ctx;
typedef struct abk_fft_plan abk_fft_plan;

/* ---- context ------------------------------------------------------------------------------ */
int abk_version(void);
const char *abk_last_error(void);
int abk_ctx_create(int device, abk_ctx **ctx);
int abk_ctx_destroy(abk_ctx *ctx);
/* `stream` is a cudaStream_t (NULL = legacy default stream). */
int abk_ctx_set_stream(abk_ctx *ctx, void *stream);
int abk_ctx_sync(abk_ctx *ctx);
/* number of kernels this library has launched on this context (bench.py "gpu_launches") */
int64_t abk_ctx_launch_count(abk_ctx *ctx);
/* Optional per-kernel timing: when enabled, every kernel launch is bracketed by a CUDA event pair
 * on the launch stream.  abk_ctx_profile_collect synchronises the stream and ADDS, per kernel id,
 * the elapsed milliseconds to ms_h[id] and the number of launches to n_h[id] (host arrays of
 * abk_kernel_count() entries), then forgets the records.  abk_kernel_name(id) names an id. */
int abk_ctx_profile_enable(abk_ctx *ctx, int on);
int abk_ctx_profile_collect(abk_ctx *ctx, double *ms_h, int64_t *n_h);
int abk_kernel_count(void);
const char *abk_kernel_name(int id);
/* Mass-assignment scheme used by the bucket / deposit entry points: 0 = TSC (analysis/tsc.py, default),
 * 1 = CIC (analysis/cic.py:13-125 `cic_serial`: same 27-cell update with weights (max(d,0), 1-|d|, max(-d,0)),
 * cell index from p = ((pos + offset) / box) * g evaluated in double). */
int abk_ctx_set_scheme(abk_ctx *ctx, int scheme);
/* Every particle weight is multiplied by `scale` when bucket records are written (abk_tsc_bucket*, abk_tsc_deposit)
 * and in abk_tsc_deposit_naive; default 1.  With scale = n^3 / N and a grid initialised to -1 the deposit itself
 * produces the normalised field rho * n^3/N - 1 of `normalize_field` (power_spectrum.py:860-901): one read+write pass
 * over the mesh less.  Records routed between GPUs (REC4 input) are scaled when they are re-bucketed, i.e. once. */
int abk_ctx_set_weight_scale(abk_ctx *ctx, double scale);
/* tuning knobs (0 = library default): tile-kernel particle capacity per pass */
int abk_ctx_set_tile_capacity(abk_ctx *ctx, int capacity);

/* ---- TSC deposit --------------------------------------------------------------------------- */

/* tsc.py:219-226 `_wrap_inplace`: one-shot periodic wrap of pos[N][3] into [0, box], in place.
 * *n_changed_dev (device int64, may be NULL) is incremented by the number of coordinates that
 * were modified, so a host caller can skip the write-back when nothing changed. */
int abk_wrap_inplace(abk_ctx *ctx, float *pos, int64_t N, double box, int64_t *n_changed_dev);

/* tsc.py:259-384 `partition_parallel`: counting sort of particles into `npart` stripes along
 * `coord`; key = min(int32(pos[coord] * f32(npart/box)), npart-1).  out_starts is int64[npart+1].
 * The order of particles inside a stripe as written by the kernel is unspecified (atomic scatter), which the
 * reference's own test allows (tests/test_tsc.py:194-208); out_index (uint32[N], may be NULL) receives the source
 * index of every output row, from which the host layer restores the reference's stable order (threads own contiguous
 * input ranges there, tsc.py:338-376).  scratch: abk_partition_scratch_bytes(). */
int abk_partition_scratch_bytes(int64_t N, int npart, size_t *bytes);
int abk_partition(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int npart, double box,
                  int coord, float *out_pos, float *out_w, int64_t *out_starts, uint32_t *out_index, void *scratch,
                  size_t scratch_bytes);

/* Particle bucketing for the deposit (replaces the x-stripe partition of tsc.py:178-188 and the
 * two-colour stripe schedule of tsc.py:229-256 `_tsc_parallel`): particles are binned by the
 * (8 x 8 x 32)-cell tile that contains the centre cell of their cloud,
 *     cell = rint((pos + offset) * f32(n/box)) mod n          (tsc.py:424-433, 387-391)
 * One call buckets one SEGMENT of up to 2^30 particles; several segments (e.g. host->device
 * chunks that arrive one after another) can be bucketed independently and deposited together.
 *   records    : out, float4[N]   (x, y, z, w) in bucket order (w = 1 when `w` is NULL)
 *   tile_starts: out, uint32[ntiles+1] exclusive scan: tile t owns records
 *                [tile_starts[t], tile_starts[t+1])
 * If `wrap` is non-zero the one-shot periodic wrap (tsc.py:219-226) is applied on the fly to the
 * values that are bucketed (the input array is not modified). */
int abk_tsc_num_tiles(int nx, int ny, int nz, int64_t *ntiles);
/* Measurement helper: float-reduction rate of this GPU in 1e9 adds/s into `buf` (device, nfloats floats, is
 * overwritten): mode 0 = coalesced rows of 32 consecutive floats per warp instruction, mode 1 = 32 scattered cells per
 * instruction.  Synchronises the stream. */
int abk_bench_red_rate(abk_ctx *ctx, float *buf, int64_t nfloats, int mode, double *gadds_per_s);
/* cells per deposit tile along x, y, z (tile id = (tx * nty + ty) * ntz + tz, n?t = ceil(n? / t?)) */
int abk_tsc_tile_shape(int *tx, int *ty, int *tz);
int abk_tsc_bucket_scratch_bytes(int64_t N, int nx, int ny, int nz, size_t *bytes);
int abk_tsc_bucket(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int nx, int ny, int nz,
                   double box, double offset, int wrap, void *records, uint32_t *tile_starts,
                   void *scratch, size_t scratch_bytes);
/* Slab mode (mesh sharded over GPUs by x-planes): only particles whose centre cell lies in planes
 * [x_lo, x_lo+nxe) (mod nx) are bucketed; tiles cover that x-range.  The number of particles that
 * fell outside is returned in *n_dropped_h (host; the call then synchronises the stream). */
int abk_tsc_bucket_slab(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int in_records, int nx, int ny,
                        int nz, double box, double offset, int wrap, int x_lo, int nxe, void *records,
                        uint32_t *tile_starts, void *scratch, size_t scratch_bytes,
                        unsigned long long *n_dropped_h);
/* in_records != 0: `pos` points to float4 (x,y,z,w) records (as produced by abk_route_particles)
 * instead of float[N][3] (+ w).
 *
 * Routing for a mesh sharded by x-planes: the owner of a particle is the rank whose plane range
 * [xsplit_h[r], xsplit_h[r+1]) contains the centre cell of its unshifted cloud,
 * rint(x * f32(nx/box)) mod nx.  Writes (x,y,z,w) records grouped by owner into records_out
 * (float4[N]; may be NULL to only count) and the per-owner counts into counts_h (host,
 * int64[nranks]); synchronises the stream.  The one-shot periodic wrap (tsc.py:219-226) is applied to
 * the routed copies when `wrap` is set. */
int abk_route_particles(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int nx, double box, int wrap,
                        int nranks, const int32_t *xsplit_h, void *records_out, int64_t *counts_h);

/* tsc.py:394-507 `_tsc_scatter` (+ :229-256): 27-point TSC deposit of bucketed particles.
 * One CTA per tile builds per-cell particle lists in shared memory, accumulates each cell's
 * clouds in registers, combines neighbouring cells conflict-free in a shared-memory tile, and
 * flushes the tile (+1-cell halo) to the grid with float reductions.  The grid is accumulated
 * into, never zeroed (tsc.py:45-50).
 *   nseg segments: records_seg_h[s] / tile_starts_seg_h[s] (host arrays of device pointers) as
 *   produced by abk_tsc_bucket with the SAME grid shape and box; seg_counts_h[s] = N of segment s
 *   (sizes the shared-memory particle capacity).  `offset` may differ from `bucket_offset`, the
 *   offset the records were bucketed with (interlacing: bucket once at offset 0, deposit at 0 and at
 *   half a cell).  The tile kernel then covers one more cell in x and y, which holds every particle
 *   whose cell moved by 0 or +1; anything else (z overflow, larger shifts) is deposited with 27
 *   direct reductions.
 *   slab == 0 : x_lo = 0, nxe = nx, the grid holds nx planes and x wraps periodically.
 *   slab != 0 : the records are those of abk_tsc_bucket_slab(x_lo, nxe); the grid holds nxe+2 planes:
 *   plane 0 is the ghost plane x_lo-1, planes 1..nxe are x_lo..x_lo+nxe-1, plane nxe+1 is the ghost
 *   plane x_lo+nxe (no wrap in x). */
int abk_tsc_deposit_tiles(abk_ctx *ctx, int nseg, const void *const *records_seg_h,
                          const uint32_t *const *tile_starts_seg_h, const int64_t *seg_counts_h, float *grid,
                          int nx, int ny, int nz, int64_t ldz, double box, double offset, double bucket_offset,
                          int slab, int x_lo, int nxe);

/* Convenience: bucket one segment and deposit it (what tsc_parallel does for device inputs).
 * scratch must hold abk_tsc_deposit_scratch_bytes(). */
int abk_tsc_deposit_scratch_bytes(int64_t N, int nx, int ny, int nz, size_t *bytes);
int abk_tsc_deposit(abk_ctx *ctx, const float *pos, const float *w, int64_t N, float *grid, int nx,
                    int ny, int nz, int64_t ldz, double box, double offset, int wrap, void *scratch,
                    size_t scratch_bytes);

/* Reference-order fallback used for validation only: one thread per particle, 27 global float
 * reductions (no bucketing).  Same arithmetic as abk_tsc_deposit. */
int abk_tsc_deposit_naive(abk_ctx *ctx, const float *pos, const float *w, int64_t N, float *grid,
                          int nx, int ny, int nz, int64_t ldz, double box, double offset, int wrap);

/* ---- real-space field ---------------------------------------------------------------------- */

/* power_spectrum.py:860-901 `normalize_field` (in place): field <- field * f32(size/tot_weight) - 1
 * over the nx*ny*nz valid cells of a (possibly padded) grid; size = size_total (the GLOBAL cell
 * count, so a slab of a sharded grid is normalised with the global constant). */
int abk_normalize_field(abk_ctx *ctx, float *grid, int64_t nx, int64_t ny, int64_t nz, int64_t ldz,
                        double size_total, double tot_weight);

/* ---- FFT ----------------------------------------------------------------------------------- */

/* power_spectrum.py:980,986,1059 `scipy.fft.rfftn`: unnormalised forward R2C transform of an
 * (nx,ny,nz) float32 grid stored in place with ldz = 2*(nz/2+1); output complex64
 * (nx,ny,nz/2+1) in the same buffer.  cuFFT plan with a caller-provided work area. */
int abk_rfft3_plan_create(abk_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, abk_fft_plan **plan,
                          size_t *work_bytes);
int abk_rfft3_exec(abk_ctx *ctx, abk_fft_plan *plan, float *grid_inplace, void *work, size_t work_bytes);
/* inverse (C2R, unnormalised) of the same layout: power_spectrum.py:645 `irfftn` (xi(r) path) */
int abk_irfft3_exec(abk_ctx *ctx, abk_fft_plan *plan, float *grid_inplace, void *work, size_t work_bytes);
int abk_fft_plan_destroy(abk_fft_plan *plan);
/* Which cuFFT the plans run on: libabk opens the library by path at first use (the CUDA toolkit's copy in preference to
 * one another package has already loaded under the same soname, see csrc/abk_fft.cu); `path` receives the name it was
 * opened by, `version` cufftGetVersion(). */
int abk_fft_backend(char *path, int path_len, int *version);

/* Slab-decomposed pieces for a mesh sharded over GPUs by x-planes:
 *   2-D R2C over (y,z) on `nplanes` local planes (in place, ldz = 2*(nz/2+1)), and
 *   1-D C2C along x for `nrows` rows of `nzc` complex values (layout [x][row][nzc], in place). */
int abk_fft_yz_plan_create(abk_ctx *ctx, int64_t nplanes, int64_t ny, int64_t nz, abk_fft_plan **plan,
                           size_t *work_bytes);
int abk_fft_x_plan_create(abk_ctx *ctx, int64_t nx, int64_t nrows, int64_t nzc, abk_fft_plan **plan,
                          size_t *work_bytes);
int abk_fft_exec_generic(abk_ctx *ctx, abk_fft_plan *plan, void *data_inplace, void *work,
                         size_t work_bytes);

/* ---- k-space ------------------------------------------------------------------------------- */

/* Description of one complex (or real) k-space mesh and the part of it this GPU holds.
 * Element (i,j,k) of the global (n,n,nzc) mesh lives at base[(i-i0)*stride_i + (j-j0)*stride_j + k]
 * for i0 <= i < i1, j0 <= j < j1, 0 <= k < nzc.  Single GPU: i0=j0=0, i1=j1=n,
 * stride_j=row length, stride_i=n*stride_j. */
typedef struct abk_kmesh {
    int32_t n;        /* global mesh size per dimension */
    int32_t nzc;      /* number of k_z values visited: n/2+1 */
    int32_t i0, i1;   /* local x-range */
    int32_t j0, j1;   /* local y-range */
    int64_t stride_i; /* in elements */
    int64_t stride_j; /* in elements */
} abk_kmesh;

/* power_spectrum.py:904-948 `shift_field_fft`, :1073-1078 `_normalize`, :1062-1070 window
 * compensation -- fused, in place on f:
 *     f <- (f + fs * exp(i*pi*(i'+j'+k)/n)) * scale      (fs != NULL; scale = 0.5/n^3)
 *     f <-  f * scale                                     (fs == NULL; scale = 1/n^3)
 *     f <-  f / ((W[i]*W[j])*W[k])                        (W != NULL; float32 table of length n)
 * i' = i (i < n/2) else i-n, likewise j'. */
int abk_field_fft_finish(abk_ctx *ctx, const abk_kmesh *mesh_h, void *f, const void *fs, const float *W,
                         float scale);
/* power_spectrum.py:904-948 `shift_field_fft` for an arbitrary shift d of the second field:
 *     f <- (f + fs * exp(i * 0.5 d (kx+ky+kz))) * scale = (f + fs * exp(i*pi*(i'+j'+k) * d/L)) * scale
 * (abk_field_fft_finish is the d = L/n case fused with the window division). */
int abk_shift_field_fft(abk_ctx *ctx, const abk_kmesh *mesh_h, void *f, const void *fs, double d_over_L, float scale);

/* power_spectrum.py:707-727 `get_raw_power`: out[t] = |f1[t]|^2, or Re(conj(f1[t]) f2[t]) when f2 != NULL,
 * over `size` complex64 elements (materialised; calc_power itself uses the fused abk_power_bin). */
int abk_raw_power(abk_ctx *ctx, const void *f1, const void *f2, float *out, int64_t size);

/* out[t] = in[t] + 0i over `size` elements: turns a real P(k) mesh into the complex input of the C2R
 * transform of power_spectrum.py:645 (`irfftn(Pk)` in pk_to_xi). */
int abk_real_to_complex(abk_ctx *ctx, const float *in, void *out, int64_t size);

/* Binning request: power_spectrum.py:150-300 `bin_kmu` fused with :707-727 `get_raw_power` and,
 * optionally, with the finishing step above (so calc_power never materialises delta(k) or P(k)).
 *
 *   value per mode:  real_in != NULL : real_in[(i,j,k)]                       (project_3d_to_poles)
 *                    f2 == NULL      : |v1|^2                                 (auto power)
 *                    else            : Re(conj(v1) * v2)                      (cross power)
 *   with v = finish(f, fs, W, scale) when `finish` is non-zero, else v = f.
 *
 *   kmag2 = f32(i'^2+j'^2+k^2); mu2 = f32(k^2)/kmag2 (0 at DC); a mode is used iff
 *   kedges2[0] <= kmag2 < kedges2[Nk]; bk = #{b in 1..Nk : kedges2[b] < kmag2},
 *   bmu = min(#{b in 1..Nmu : muedges2[b] < mu2}, Nmu-1); multiplicity 1 for k==0, else 2.
 *
 * Outputs are RAW sums, ACCUMULATED into (the caller zeroes them; several GPUs add theirs with an
 * all-reduce): counts u64[Nk*Nmu], sum_p f64[Nk*Nmu] (sum of mult*value), sum_k f64[Nk*Nmu]
 * (sum of mult*sqrt(kmag2); multiply by dk on the host), sum_poles f64[Np*Nk]
 * (sum of mult*value*(2l+1)P_l(mu)); the l=0 row is left untouched (the host sets it from sum_p,
 * power_spectrum.py:282-284).
 * pole_coef: device float32[Np][ABK_POLE_NCOEF], (2l+1)P_l as a polynomial in mu=sqrt(mu2) (host-built
 * from power_spectrum.py:121-147); pole_ell: the Np multipole orders, stored in the request. */
typedef struct abk_bin_request {
    abk_kmesh mesh;
    const void *f1, *f1s, *f2, *f2s; /* complex64 meshes (fs: half-cell-shifted grids) */
    const float *real_in;            /* float32 mesh instead of f1 */
    const float *W;                  /* float32[n] or NULL */
    float scale;
    int32_t finish;
    const float *kedges2; /* float32[Nk+1], device */
    const float *muedges2; /* float32[Nmu+1], device */
    int32_t Nk, Nmu, Np;
    const float *pole_coef; /* device */
    int32_t pole_ell[ABK_MAX_POLES];
    unsigned long long *counts;
    double *sum_p, *sum_k, *sum_poles;
    /* optional device scratch of abk_power_bin_scratch_bytes(): lets the kernel spread its
     * reductions over replicated sum tables (less same-address contention); NULL = reduce straight
     * into the outputs */
    void *scratch;
    size_t scratch_bytes;
    /* non-zero if the host verified W[n-a] == W[a] for all a (or W is NULL): enables the kernel that
     * bins the mirror modes (+-i', +-j', k) together */
    int32_t w_symmetric;
} abk_bin_request;

int abk_power_bin_scratch_bytes(int Nk, int Nmu, int Np, size_t *bytes);
int abk_power_bin(abk_ctx *ctx, const abk_bin_request *req_h);

/* ---- k-space field helpers of the reference's ZCV modules ------------------------------------- */

/* power_spectrum.py:577-617 `get_delta_mu2`: out = delta * mu^2, complex64 (n,n,n/2+1), mu^2 = k_z^2/|k|^2 (0 at DC) */
int abk_delta_mu2(abk_ctx *ctx, const void *delta, void *out, int n);
/* power_spectrum.py:539-574 `get_smoothing`: out[i,j,k] = exp(-|k|^2 R^2 / 2), float32 (n,n,n/2+1) */
int abk_smoothing(abk_ctx *ctx, float *out, int n, double L, double R);
/* power_spectrum.py:450-536 `expand_poles_to_3d`: out[i,j,k] = sum_l interp(k_ell, P_ell[l])(|k|) * P_l(mu);
 * k_ell float32[Nk] (uniform), P_ell float32[Np][Nk], coef float32[Np][ABK_POLE_NCOEF] = P_l as a polynomial
 * in mu (host-built, WITHOUT the 2l+1 factor), all on the device; poles_h on the host. */
int abk_expand_poles_to_3d(abk_ctx *ctx, float *out, int n, double L, const float *k_ell, const float *P_ell, int Nk,
                           const int32_t *poles_h, int Np, const float *coef);

/* bin_kppi (analysis/power_spectrum.py:303-412): count and sum the modes of an (n, n, >= n/2+1) mesh in
 * (k_perp, pi) bins.  weights: device, float32 (weights_f64 = 0) or float64 (1), row stride ldz elements
 * (n/2+1 for a half-spectrum, n for a real-space mesh binned with fourier=False).  kedges2 [Nk+1] and
 * piedges2 [Npi+1]: device float64, the squared edges in units of dk^2 AFTER rounding to the compute dtype
 * (:365-366).  kperp_f32 != 0 rounds i'^2 + j'^2 to float32 before the comparison (dtype=float32).
 * Bins are (lo, hi]; k_perp^2 < kedges2[0] is skipped; a row i stops at its first j with
 * k_perp^2 >= kedges2[Nk] (the reference's `break`, :379-380); kz^2 >= piedges2[Npi] is dropped; a mode
 * counts once on the k = 0 plane and twice elsewhere.  counts u64 [Nk*Npi] and sum_w f64 [Nk*Npi] are
 * accumulated into (the caller zeroes them and divides). */
int abk_bin_kppi(abk_ctx *ctx, const void *weights, int weights_f64, int n, int64_t ldz, const double *kedges2, int Nk,
                 const double *piedges2, int Npi, int kperp_f32, unsigned long long *counts, double *sum_w);

/* ---- device-side particle ingest --------------------------------------- */

/* `unpack_rvint` (abacusnbody/data/bitpacked.py:29-120): intdata int32[N][3] on the device; every int32 holds 20
 * signed bits of position (units of boxsize/1e6) above 12 bits of velocity (offset 2048, units of 6000/2048 km/s).
 * posout / velout: device [N][3] float32 (out_f64 = 0) or float64 (1); either may be NULL (not unpacked).
 * Values are bit-identical to the reference (float64 product, one rounding to the output type). */
int abk_unpack_rvint(abk_ctx *ctx, const int32_t *intdata, int64_t N, double boxsize, void *posout, void *velout,
                     int out_f64);

/* `unpack_pids` (abacusnbody/data/bitpacked.py:123-311): fields of the packed 64-bit PID/aux word, packed uint64[N] on
 * the device.  Any output may be NULL: pid int64[N] (the three 15-bit Lagrangian indices in place), lagr_pos [N][3]
 * float32/float64 (index * T(box/ppd) - T(box/2), float64 arithmetic rounded once), lagr_idx int16[N][3], tagged
 * uint8[N] (bit 48), density [N] float32/float64 (bits 49-58, squared). */
int abk_unpack_pids(abk_ctx *ctx, const uint64_t *packed, int64_t N, double box, int64_t ppd, int64_t *pid, void *lagr_pos,
                    int16_t *lagr_idx, uint8_t *tagged, void *density, int out_f64);

/* `unpack_pack9` (abacusnbody/data/pack9.py:16-123) in two calls, because the number of particle records is only
 * known after the cell headers (records whose first byte is 0xFF) have been counted:
 *   abk_pack9_count   counts headers per 256-record block and scans the counts into `scratch`
 *                     (abk_pack9_scratch_bytes(nrec), 256-byte aligned); SYNCHRONISES the stream and returns the
 *                     number of headers in *nheaders_h (host).  Particles = nrec - headers.
 *   abk_pack9_decode  decodes the headers into hdr_tab (device, nheaders * 5 * sizeof(out type) bytes) and the
 *                     particles into posout / velout (device [nrec - nheaders][3], either may be NULL), in stream
 *                     order.  `scratch` must still hold the result of abk_pack9_count for the same data.
 * data: device uint8[nrec][9], 4-byte aligned.  Particle records that precede the first header decode to NaN, as in
 * the reference.  All roundings follow the reference's (non-fastmath) Numba kernel for the chosen output type. */
int abk_pack9_scratch_bytes(int64_t nrec, size_t *bytes);
int abk_pack9_count(abk_ctx *ctx, const uint8_t *data, int64_t nrec, void *scratch, size_t scratch_bytes,
                    int64_t *nheaders_h);
int abk_pack9_decode(abk_ctx *ctx, const uint8_t *data, int64_t nrec, double boxsize, double velzspace_to_kms,
                     const void *scratch, void *hdr_tab, int64_t nheaders, void *posout, void *velout, int out_f64);

/* ---- multi-GPU helpers (x-slab sharded mesh) ------------------------------------------------ */

/* dst[i] += src[i] over an (nplanes, ny, nz) region of padded grids (ghost-plane accumulation) */
int abk_add_planes(abk_ctx *ctx, float *dst, const float *src, int64_t nplanes, int64_t ny, int64_t nz,
                   int64_t ldz);
/* Slab->pencil transpose, pack side: from the local slab [nxl][ny][nzc] (complex64) gather, for
 * every destination rank r, the block [nxl][j in rank r's y-range][nzc] contiguously into
 * sendbuf at element offset nxl*jsplit_h[r]*nzc.  y-ranges are given by jsplit_h[nranks+1] (host).
 * No unpack kernel is needed: if rank r's block is received at element offset
 * isplit[r]*nyl*nzc, the receive buffer IS the pencil layout [nx][nyl][nzc]. */
int abk_transpose_pack(abk_ctx *ctx, const void *slab, void *sendbuf, int64_t nxl, int64_t ny, int64_t nzc,
                       int nranks, const int64_t *jsplit_h);
/* Fused pack + transfer over NVLink peer memory: row (x, j) of the local slab [nxl][ny][nzc] is stored directly
 * at its final position [(x_lo + x)][j - jsplit[r]][.] of the pencil buffer of the rank r that owns j.
 * peer_pencils_h[r] (host array) is the peer-mapped device pointer of rank r's pencil buffer (e.g. from
 * torch symmetric memory).  The caller synchronises the ranks before (buffer free) and after (writes landed). */
int abk_transpose_scatter_p2p(abk_ctx *ctx, const void *slab, void *const *peer_pencils_h, int64_t nxl, int64_t ny,
                              int64_t nzc, int nranks, const int64_t *jsplit_h, int64_t x_lo);

/* ---- float64 --------------------------------------------------------------------
 * The reference's TSC computes in the dtype of the positions and accumulates in the dtype of the grid (tsc.py:400,
 * :471-507); calc_power(dtype=float64) honours the dtype on its non-interlaced branch (power_spectrum.py:1053-1069).
 * float32 / float32 is the fast path (abk_tsc_deposit); every other combination takes these entry points:
 * one thread per particle, 27 global reductions in the grid's dtype, arithmetic in the positions' dtype.
 *   pos  : float32 or float64 [N][3] (pos_f64), w: NULL or float32 / float64 [N] (w_f64), grid: float32 / float64 [nx][ny][ldz] */
int abk_tsc_deposit_typed(abk_ctx *ctx, const void *pos, int pos_f64, const void *w, int w_f64, int64_t N, void *grid,
                          int grid_f64, int nx, int ny, int nz, int64_t ldz, double box, double offset, int wrap);
/* normalize_field in float64: grid = grid * (size_total / tot_weight) - 1 (power_spectrum.py:860-901) */
int abk_normalize_field_f64(abk_ctx *ctx, double *grid, int64_t nx, int64_t ny, int64_t nz, int64_t ldz, double size_total,
                            double tot_weight);
/* in-place D2Z transform of a padded float64 grid [nx][ny][2(nz/2+1)] -> complex128 [nx][ny][nz/2+1] (scipy.fft.rfftn
 * of a float64 field, power_spectrum.py:1059); synchronises the stream */
int abk_rfft3_f64(abk_ctx *ctx, double *grid_inplace, int64_t nx, int64_t ny, int64_t nz);
/* power_spectrum.py:1058-1069 + :707-727 on complex128 spectra: f *= inv_size, f /= (W_i W_j) W_k (W may be NULL),
 * out = |f1|^2 or Re(conj(f1) f2) as the float32 mesh [n][n][n/2+1] that abk_power_bin reads (real_in) */
int abk_power_from_f64(abk_ctx *ctx, const void *f1, const void *f2, const float *W, int n, double inv_size, float *out);

#ifdef __cplusplus
}
#endif
#endif /* ABK_H */
