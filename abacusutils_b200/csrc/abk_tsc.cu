// TSC mass assignment on B200 (sm_100a).
//
// Replaces, with a different algorithm, the reference's Numba kernels
//   _wrap_inplace        analysis/tsc.py:219-226
//   partition_parallel   analysis/tsc.py:259-384
//   _tsc_parallel        analysis/tsc.py:229-256   (two-colour x-stripe schedule)
//   _tsc_scatter         analysis/tsc.py:394-507   (27 read-modify-writes per particle)
//
// Design:
//   A. bucket particles by the (8 x 8 x 30)-cell tile of their cloud's centre cell: histogram
//      (one 4-byte reduction per particle), scan, scatter of 16-byte (x,y,z,w) records.  The open
//      write frontier is one 128-byte line per tile; it stays (mostly) resident in the 126 MB L2.
//   B. one CTA per tile ("walk" kernel): per-cell particle lists are built in shared memory with one
//      integer exchange per particle; then warp = y-row, lane = z-cell (+ one halo lane either side),
//      x walked serially: every lane sums the 27 stencil weights of ITS cell's particles in registers
//      (a rolling 3-plane window along x, packed FFMA2 arithmetic), neighbouring lanes are combined with
//      two shuffles per row, and each finished row goes straight to the grid as ONE coalesced 32-float
//      reduction (REDG.ADD.F32).  No shared-memory output tile, no block barrier inside the walk, no
//      float shared-memory atomics (a CAS loop on sm_100a).
//   C. interlacing: the half-cell-shifted deposit reuses the records bucketed for the unshifted grid
//      (template EXT: one more cell per tile in x, y and z).  CIC (analysis/cic.py) is the same update
//      with other weights (template CIC).
#include "abk_common.cuh"

namespace {

struct TscParams {
    float inv_hx, inv_hy, inv_hz;  // f32(g/box), tsc.py:408-411
    float off;                     // f32(offset), tsc.py:413: the offset of THIS deposit / bucketing
    double box;
    int nx, ny, nz;                // global grid
    int nxe;                       // x-extent of the tiled region (== nx single GPU; slab mode: local planes)
    int x_lo;                      // first global x-plane of the tiled region
    int nty, ntz;
    int wrap;
    int cic;                       // 0: TSC (tsc.py), 1: CIC (cic.py:13-125)
    double gx_d, gy_d, gz_d;       // CIC works in double: p = (x / box) * g
    float wscale;                  // multiplies every weight as the bucket records are written (1 unless the caller
                                   // folds the field normalisation into the deposit, abk_ctx_set_weight_scale)
    int64_t plane_bytes, wrap_bytes;  // tile deposit: bytes of one x-plane of the grid (ny * ldz * 4) and of nx of them
};

// tsc.py:219-226: one-shot wrap; compare against the double box, store float32
__device__ __forceinline__ float wrap_coord(float v, double box)
{
    if ((double)v >= box) return (float)((double)v - box);
    if (v < 0.0f) return (float)((double)v + box);
    return v;
}

// tsc.py:424-440: p = (x + off) * inv_h; i = round-half-even(p); d = f32(i) - p
__device__ __forceinline__ void cell_of(float x, float off, float inv_h, int n, int &cell, float &d)
{
    const float p = __fmul_rn(__fadd_rn(x, off), inv_h);
    const float r = rintf(p);
    d = r - p;
    cell = abk_wrap_cell((int)r, n);
}

// cic.py:29-42: p = ((pos + d) / boxsize) * g in double (pos + d is a float32 sum); i = round-half-even(p)
__device__ __forceinline__ void cell_of_cic(float x, float off, double box, double g, int n, int &cell, float &d)
{
    const double p = ((double)__fadd_rn(x, off) / box) * g;
    const double r = rint(p);
    d = (float)(r - p);
    cell = abk_wrap_cell((int)r, n);
}

__device__ __forceinline__ void cells_of(const TscParams &P, float x, float y, float z, int &cx, int &cy, int &cz,
                                         float &dx, float &dy, float &dz)
{
    if (P.cic) {
        cell_of_cic(x, P.off, P.box, P.gx_d, P.nx, cx, dx);
        cell_of_cic(y, P.off, P.box, P.gy_d, P.ny, cy, dy);
        cell_of_cic(z, P.off, P.box, P.gz_d, P.nz, cz, dz);
    } else {
        cell_of(x, P.off, P.inv_hx, P.nx, cx, dx);
        cell_of(y, P.off, P.inv_hy, P.ny, cy, dy);
        cell_of(z, P.off, P.inv_hz, P.nz, cz, dz);
    }
}

// Cell index relative to a tile origin (global cell index `origin`), fast path for the common case that
// the rounded position already lies inside [origin, origin + T); periodic images take the slow path.
template <bool CIC>
__device__ __forceinline__ void local_cell(float x, float off, float inv_h, double box, double g, int n, int origin, int T,
                                           int &l, float &d)
{
    int c;
    if (CIC) {
        const double p = ((double)__fadd_rn(x, off) / box) * g;
        const double r = rint(p);
        d = (float)(r - p);
        c = (int)r;
    } else {
        const float p = __fmul_rn(__fadd_rn(x, off), inv_h);
        const float r = rintf(p);
        d = r - p;
        c = (int)r;
    }
    l = c - origin;
    if ((unsigned)l >= (unsigned)T) {
        l = abk_wrap_cell(c, n) - origin;
        if (l < 0) l += n;
    }
}

__device__ __forceinline__ bool tile_of(const TscParams &P, float x, float y, float z, uint32_t &tile)
{
    int cx, cy, cz;
    float d, d2, d3;
    cells_of(P, x, y, z, cx, cy, cz, d, d2, d3);
    int lx = cx - P.x_lo;
    if (lx < 0) lx += P.nx;
    if (lx >= P.nxe) return false;  // not owned by this slab
    tile = ((uint32_t)(lx / ABK_TX) * P.nty + (uint32_t)(cy / ABK_TY)) * P.ntz + (uint32_t)(cz / ABK_TZ);
    return true;
}

// Four particles per thread: 3 x 128-bit loads of positions (AoS float[N][3]), then all four tile
// atomics are issued back to back (independent, so their L2 round trips overlap), then the stores.
template <bool SCATTER, bool REC4>
__global__ void __launch_bounds__(256) tsc_bucket_kernel(const float *__restrict__ pos, const float *__restrict__ w,
                                                         int64_t N, TscParams P, uint32_t *__restrict__ counts,
                                                         float4 *__restrict__ records, int vec_ok,
                                                         unsigned long long *__restrict__ n_dropped)
{
    const int64_t ngroups = (N + 3) / 4;
#if defined(__CUDA_ARCH__)
    uint64_t keep_policy = 0;
    if (SCATTER) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
#endif
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t base = g * 4;
        float c[12];
        float wv[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        const int cnt = (int)min((int64_t)4, N - base);
        if (REC4) {  // input already is (x,y,z,w) records (routed particles of a sharded run)
            const float4 *p4 = reinterpret_cast<const float4 *>(pos) + base;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float4 r = make_float4(0.f, 0.f, 0.f, 1.f);
                if (q < cnt) r = __ldcs(p4 + q);
                c[3 * q] = r.x; c[3 * q + 1] = r.y; c[3 * q + 2] = r.z; wv[q] = r.w;
            }
        } else if (vec_ok && cnt == 4) {
            const float4 *p4 = reinterpret_cast<const float4 *>(pos + 3 * base);
            const float4 a = __ldcs(p4), b = __ldcs(p4 + 1), d = __ldcs(p4 + 2);
            c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
            c[8] = d.x; c[9] = d.y; c[10] = d.z; c[11] = d.w;
        } else {
#pragma unroll
            for (int q = 0; q < 12; q++) c[q] = (q < 3 * cnt) ? __ldcs(pos + 3 * base + q) : 0.0f;
        }
        if (!REC4 && SCATTER && w) {
            if (vec_ok && cnt == 4) {
                const float4 t = __ldcs(reinterpret_cast<const float4 *>(w + base));
                wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) if (q < cnt) wv[q] = __ldcs(w + base + q);
            }
        }
        uint32_t tile[4];
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (P.wrap) {
                c[3 * q] = wrap_coord(c[3 * q], P.box);
                c[3 * q + 1] = wrap_coord(c[3 * q + 1], P.box);
                c[3 * q + 2] = wrap_coord(c[3 * q + 2], P.box);
            }
            ok[q] = (q < cnt) && tile_of(P, c[3 * q], c[3 * q + 1], c[3 * q + 2], tile[q]);
        }
        if (!SCATTER) {
            unsigned dropped = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (ok[q]) atomicAdd(&counts[tile[q]], 1u);
                else if (q < cnt) dropped++;
            }
            if (dropped) atomicAdd(n_dropped, (unsigned long long)dropped);
        } else {
            // cursors count DOWN from the inclusive scan, so they end as the exclusive scan
            uint32_t slot[4];
#pragma unroll
            for (int q = 0; q < 4; q++) slot[q] = ok[q] ? atomicSub(&counts[tile[q]], 1u) - 1u : 0u;
            // the record stores ask the L2 to keep their lines (evict_last): a tile's open 128-byte line is written eight
            // times before it is complete, and the streamed positions (evict_first loads) should not push it out in between
            // -- 26.7 vs 27.7 ms per 1e9 particles at the segment size of config 3 (scripts/micro/split_micro.cu)
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (ok[q]) {
                    const float4 rec = make_float4(c[3 * q], c[3 * q + 1], c[3 * q + 2], wv[q] * P.wscale);
#if defined(__CUDA_ARCH__)
                    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(records + slot[q]), "f"(rec.x), "f"(rec.y),
                                 "f"(rec.z), "f"(rec.w), "l"(keep_policy)
                                 : "memory");
#else
                    records[slot[q]] = rec;
#endif
                }
        }
    }
}

constexpr uint32_t NIL = 0xffffffffu;

struct SegList {
    int nseg;
    const float4 *rec[ABK_MAX_SEGMENTS];
    const uint32_t *starts[ABK_MAX_SEGMENTS];
};

// tsc.py:442-451: the three 1-D TSC weights for cells i-1, i, i+1 given d = i - p
__device__ __forceinline__ void tsc_w(float d, float &wm, float &w0, float &wp)
{
    const float a = 0.5f + d, b = 0.5f - d;
    wm = 0.5f * a * a;
    w0 = 0.75f - d * d;
    wp = 0.5f * b * b;
}

// cic.py:43-67: weights of cells i-1, i, i+1 for d = i - p: (max(d,0), 1-|d|, max(-d,0))
__device__ __forceinline__ void cic_w(float d, float &wm, float &w0, float &wp)
{
    wm = fmaxf(d, 0.0f);
    w0 = 1.0f - fabsf(d);
    wp = fmaxf(-d, 0.0f);
}

template <bool CIC>
__device__ __forceinline__ void mas_w(float d, float &wm, float &w0, float &wp)
{
    if (CIC) cic_w(d, wm, w0, wp);
    else tsc_w(d, wm, w0, wp);
}

// 27 global reductions for a particle whose centre cell lies outside the cell domain of the tile it
// was bucketed in (only possible when the deposit offset differs from the bucketing offset).
// (the mesh shape comes by value: a reference to the kernel's parameter struct would force a copy of it into local memory,
// and every later use of a field -- also in the hot loops -- would become a local load)
struct MeshShape { int nx, ny, nz, x_lo, cic; };
__device__ __noinline__ void deposit_direct(float *__restrict__ grid, const MeshShape P, int64_t ldz, int slab,
                                            int cx, int cy, int cz, float dx, float dy, float dz, float W)
{
    const int64_t sx = (int64_t)P.ny * ldz;
    float wx[3], wy[3], wz[3];
    if (P.cic) {
        cic_w(dx, wx[0], wx[1], wx[2]); cic_w(dy, wy[0], wy[1], wy[2]); cic_w(dz, wz[0], wz[1], wz[2]);
    } else {
        tsc_w(dx, wx[0], wx[1], wx[2]); tsc_w(dy, wy[0], wy[1], wy[2]); tsc_w(dz, wz[0], wz[1], wz[2]);
    }
    for (int a = 0; a < 3; a++) {
        int64_t gx;
        if (slab) {
            int lx = cx - P.x_lo;
            if (lx < 0) lx += P.nx;
            gx = lx + a;  // plane 0 of a slab grid is the ghost plane x_lo-1
        } else {
            gx = abk_wrap_cell(cx + a - 1, P.nx);
        }
        for (int b = 0; b < 3; b++) {
            const int gy = abk_wrap_cell(cy + b - 1, P.ny);
            for (int c = 0; c < 3; c++) {
                const int gz = abk_wrap_cell(cz + c - 1, P.nz);
                atomicAdd(grid + gx * sx + (int64_t)gy * ldz + gz, wx[a] * wy[b] * wz[c] * W);
            }
        }
    }
}

// ==============================================================================================
// Tile deposit ("walk" kernel).  One CTA per tile of 8 x 8 x 30 cells.
//   phase 1 (lane <-> particle): the tile's records arrive in shared memory by bulk copy (cp.async.bulk + mbarrier), are
//            converted in place to (dx, dy, dz, W) and threaded onto per-cell lists with one shared-memory exchange
//            each (ATOMS.EXCH runs at ~1.3e12 lane-ops/s on B200, scripts/micro/deposit_micro.cu);
//   phase 2 (lane <-> cell): warp = y-row of the tile, lanes 1..30 = its z-cells, lanes 0 and 31 = the z-halo cells;
//            x walked serially.  Every lane sums the clouds of the particles of ITS cell into a rolling window of three
//            x-planes held in registers, stored as packed float pairs so the 27 multiply-adds of a particle are
//            12 FFMA2 + 3 FFMA with scalar-broadcast operands.  A finished plane is combined with the neighbouring lanes
//            by two rotating shuffles per row (the halo lanes hold no particles, so the wrap-around terms are zero) and
//            added STRAIGHT to the grid: three coalesced 32-float reductions per x-step, halo cells included.  There is
//            no shared-memory output tile and no block barrier inside the walk: the warps of a CTA run independently, so
//            an x-step costs the longest of the warp's 30 lists, not of the CTA's 240.
// EXT = 1 (records bucketed at offset 0, deposit at +half a cell): the centre cell is the bucketed one or its +1
//            neighbour, so the tile's cell domain grows to 9 x 9 x 31; lane 31 then owns a cell and its upper z-term
//            goes through a small per-warp stash that is reduced at the end of the walk.
// Plane layout (9 floats): p[b] = (S[b][z-1], S[b][z+1]) for the three rows b = y-1, y, y+1; q = (S[y-1][z], S[y+1][z]);
// s = S[y][z].  TSC weights come as natural pairs (w-, w+) from one packed square, w0 = 0.75 - d^2.
struct Plane {
    float2 p[3];
    float2 q;
    float s;
};

__device__ __forceinline__ void plane_zero(Plane &P)
{
    P.p[0] = P.p[1] = P.p[2] = P.q = make_float2(0.0f, 0.0f);
    P.s = 0.0f;
}

// P += xa * (cloud of one particle in the y-z plane): T[b] = (y_b z-, y_b z+), U = (y- z0, y+ z0), u0 = y0 z0
__device__ __forceinline__ void plane_fma(Plane &P, float xa, const float2 (&T)[3], float2 U, float u0)
{
    const float2 xx = make_float2(xa, xa);
    P.p[0] = __ffma2_rn(T[0], xx, P.p[0]);
    P.p[1] = __ffma2_rn(T[1], xx, P.p[1]);
    P.p[2] = __ffma2_rn(T[2], xx, P.p[2]);
    P.q = __ffma2_rn(U, xx, P.q);
    P.s = fmaf(u0, xa, P.s);
}

// (w-, w+) and w0 of tsc.py:442-451 / cic.py:43-67 for d = i - p, optionally scaled by W
template <bool CIC>
__device__ __forceinline__ void mas_pair(float d, float W, float2 &wmp, float &w0)
{
    if (CIC) {
        wmp = make_float2(fmaxf(d, 0.0f) * W, fmaxf(-d, 0.0f) * W);
        w0 = (1.0f - fabsf(d)) * W;
    } else {
        const float2 a = __fadd2_rn(make_float2(d, -d), make_float2(0.5f, 0.5f));
        const float h = 0.5f * W;
        wmp = __fmul2_rn(__fmul2_rn(a, a), make_float2(h, h));
        w0 = fmaf(-d, d, 0.75f) * W;
    }
}

template <bool CIC>
__device__ __forceinline__ void accumulate_particle(const float4 r, Plane &A, Plane &B, Plane &C)
{
    if (CIC) {
        float2 X2, Y2, Z2;
        float x0, y0, z0;
        mas_pair<true>(r.x, 1.0f, X2, x0);
        mas_pair<true>(r.y, r.w, Y2, y0);
        mas_pair<true>(r.z, 1.0f, Z2, z0);
        float2 T[3];
        T[0] = __fmul2_rn(Z2, make_float2(Y2.x, Y2.x));
        T[1] = __fmul2_rn(Z2, make_float2(y0, y0));
        T[2] = __fmul2_rn(Z2, make_float2(Y2.y, Y2.y));
        const float2 U = __fmul2_rn(Y2, make_float2(z0, z0));
        const float u0 = y0 * z0;
        plane_fma(A, X2.x, T, U, u0);
        plane_fma(B, x0, T, U, u0);
        plane_fma(C, X2.y, T, U, u0);
        return;
    }
    // TSC (tsc.py:442-451): w-+ = 0.5 (0.5 +- d)^2 = (c +- d/sqrt2)^2 with c = 0.5/sqrt2 -- one packed FMA and one packed
    // square per pair, no separate halving; x and y share the packed operations ((dx, dy) is a register pair of the
    // record as loaded), the particle weight rides on the z factors.
    constexpr float K = 0.70710678118654752f, CC = 0.35355339059327376f;
    const float2 dxy = make_float2(r.x, r.y);
    const float2 am = __ffma2_rn(dxy, make_float2(K, K), make_float2(CC, CC));
    const float2 ap = __ffma2_rn(dxy, make_float2(-K, -K), make_float2(CC, CC));
    const float2 Sm = __fmul2_rn(am, am);                                              // (x-, y-)
    const float2 Sp = __fmul2_rn(ap, ap);                                              // (x+, y+)
    const float2 S0 = __ffma2_rn(make_float2(-r.x, -r.y), dxy, make_float2(0.75f, 0.75f));  // (x0, y0)
    const float zm = fmaf(r.z, K, CC), zp = fmaf(r.z, -K, CC);
    const float2 Z2 = __fmul2_rn(__fmul2_rn(make_float2(zm, zp), make_float2(zm, zp)), make_float2(r.w, r.w));  // (z-, z+) W
    const float z0 = fmaf(-r.z, r.z, 0.75f) * r.w;
    float2 T[3];
    T[0] = __fmul2_rn(Z2, make_float2(Sm.y, Sm.y));
    T[1] = __fmul2_rn(Z2, make_float2(S0.y, S0.y));
    T[2] = __fmul2_rn(Z2, make_float2(Sp.y, Sp.y));
    const float2 U = make_float2(Sm.y * z0, Sp.y * z0);
    const float u0 = S0.y * z0;
    plane_fma(A, Sm.x, T, U, u0);
    plane_fma(B, S0.x, T, U, u0);
    plane_fma(C, Sp.x, T, U, u0);
}

template <int EXT>
struct WalkDom {
    static constexpr int NXC = ABK_TX + EXT, NYC = ABK_TY + EXT, NZC = ABK_TZ + EXT;  // cells with particle lists
    static constexpr int NCELL = NXC * NYC * 32;                  // list heads: [x][y][lane], lane = local z + 1
    static constexpr int NW = NYC, NT = NW * 32;                  // one warp per y-row
    static constexpr int NPL = NXC + 2;                           // x-planes a row walk emits (1-cell halo either side)
    static constexpr int STASH = EXT ? NPL * 3 : 0;               // upper z-halo values of one walk: [plane][row]
};
static_assert(ABK_TZ + 1 + 1 <= 32, "a tile's z-cells and its two halo cells must fit one warp");

static size_t walk_smem_bytes(int cap, int ext)
{
    const size_t ncell = ext ? WalkDom<1>::NCELL : WalkDom<0>::NCELL;
    const size_t stash = ext ? (size_t)WalkDom<1>::NW * WalkDom<1>::STASH : 0;
    return abk_align_up((ncell + stash) * 4, 16) + (size_t)cap * 18 + 64;
}

// ---- bulk global -> shared copies completing on an mbarrier (cp.async.bulk; SASS: UBLKCP + SYNCS) ---------------
// One elected thread stages the tile's records: no register round trip, no LSU issue slots, and the tile's 256+ threads
// are free to initialise the list heads meanwhile.  The CPU emulator build (tests/emu) copies with a plain loop.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// i in [-n, 2n): one conditional step instead of abk_wrap_cell's general modulo (plane / row indices next to a tile)
__device__ __forceinline__ int wrap_near(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

template <int EXT, bool CIC>
__global__ void __launch_bounds__(WalkDom<EXT>::NT, EXT ? 3 : 4)
tsc_tile_walk_kernel(SegList segs, float *__restrict__ grid, TscParams P, int64_t ldz, int cap, int slab)
{
    using D = WalkDom<EXT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int REC_OFF_BYTES = ((D::NCELL + D::NW * D::STASH) * 4 + 15) / 16 * 16;
    uint32_t *head = reinterpret_cast<uint32_t *>(smem_raw);
    float *stash_all = reinterpret_cast<float *>(head + D::NCELL);
    float4 *srec = reinterpret_cast<float4 *>(smem_raw + REC_OFF_BYTES);
    uint16_t *next16 = reinterpret_cast<uint16_t *>(srec + cap);
    __shared__ uint32_t seg_beg[ABK_MAX_SEGMENTS], seg_off[ABK_MAX_SEGMENTS + 1];
    __shared__ __align__(8) uint64_t bar;
    // particles whose cell lies outside the tile's cell domain (only possible when the deposit offset is not the
    // bucketing offset or that plus half a cell): queued here and deposited warp-cooperatively, one lane per stencil point
    constexpr int MAX_OVF = 128;
    __shared__ uint16_t ovf_v[MAX_OVF], ovf_xyz[MAX_OVF];
    __shared__ unsigned ovf_cnt;

    const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
    // 3-D grid (tz, ty, tx): the tile's coordinates come from the block index, not from three integer divisions per thread
    const uint32_t tz = blockIdx.x, ty = blockIdx.y, tx = blockIdx.z;
    const uint32_t tile = (tx * gridDim.y + ty) * gridDim.x + tz;
    // the tile's record range in every segment: loaded by one lane per segment (the loads overlap), then prefix-summed
    if (tid < 32) {
        uint32_t b = 0, cnt = 0;
        if (tid < segs.nseg) {
            b = segs.starts[tid][tile];
            cnt = segs.starts[tid][tile + 1] - b;
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < ABK_MAX_SEGMENTS; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (tid < segs.nseg) { seg_beg[tid] = b; seg_off[tid] = inc - cnt; }
        if (tid == segs.nseg - 1) seg_off[segs.nseg] = inc;
#if defined(__CUDA_ARCH__)
        if (tid == 0) mbar_init(&bar, 1);
#endif
    }
    __syncthreads();
    const uint32_t total = seg_off[segs.nseg];
    if (total == 0) return;

    const int x0 = tx * ABK_TX, y0 = ty * ABK_TY, z0 = tz * ABK_TZ;  // x0 relative to x_lo
    const int64_t sx = (int64_t)P.ny * ldz;

    // plane-independent pieces of this warp's output addresses: rows y0+wy-1 .. y0+wy+1, column z0-1+lane.  Rows or
    // columns beyond the mesh (ragged last tile, meshes smaller than a tile) only ever carry zeros, which are not written.

    const int ldz32 = (int)ldz;
    const int gy0 = abk_wrap_cell(y0 + wy - 1, P.ny), gy1 = wrap_near(gy0 + 1, P.ny), gy2 = wrap_near(gy1 + 1, P.ny);
    const int rowo0 = gy0 * ldz32 + abk_wrap_cell(z0 - 1 + lane, P.nz);
    const int d01 = (gy1 - gy0) * ldz32, d12 = (gy2 - gy1) * ldz32;  // warp-uniform row steps (ldz, or back to row 0)
    const int gx_first = slab ? x0 : wrap_near(x0 - 1, P.nx);        // global x of plane 0 (local x = -1)
    const int d01b = d01 * 4, d12b = d12 * 4;                         // the same row steps in bytes
    float *stash = stash_all + wy * D::STASH;
    uint32_t parity = 0;

    for (uint32_t chunk0 = 0; chunk0 < total; chunk0 += cap) {
        const int m = (int)min((uint32_t)cap, total - chunk0);
        if (chunk0) __syncthreads();  // the previous pass' walks are done with the lists and records
        // ---- stage the raw records of this pass: one bulk copy per segment that overlaps [chunk0, chunk0 + m) ----
#if defined(__CUDA_ARCH__)
        if (tid < 32) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&bar, (uint32_t)m * 16u);
            }
            __syncwarp();
            if (tid < segs.nseg) {
                const int s = tid;
                const uint32_t lo = max(seg_off[s], chunk0), hi = min(seg_off[s + 1], chunk0 + (uint32_t)m);
                if (lo < hi) bulk_g2s(srec + (lo - chunk0), segs.rec[s] + seg_beg[s] + (lo - seg_off[s]), (hi - lo) * 16u, &bar);
            }
        }
#else
        for (int s = 0; s < segs.nseg; s++) {
            const uint32_t lo = max(seg_off[s], chunk0), hi = min(seg_off[s + 1], chunk0 + (uint32_t)m);
            for (uint32_t u = lo + tid; u < hi; u += D::NT) srec[u - chunk0] = segs.rec[s][seg_beg[s] + (u - seg_off[s])];
        }
#endif
        for (int c = tid; c < D::NCELL; c += D::NT) head[c] = NIL;
        if (tid == 0) ovf_cnt = 0;
        __syncthreads();
#if defined(__CUDA_ARCH__)
        mbar_wait(&bar, parity);
        parity ^= 1u;
#endif
        // ---- phase 1, lane <-> particle: (dx, dy, dz, W) records in place, per-cell lists --------------------------
        const int ox_g = P.x_lo + x0;  // global cell index of the tile origin in x (slab: may exceed nx, handled by wrap)
        const int ox_w = ox_g >= P.nx ? ox_g - P.nx : ox_g;
        for (int v = tid; v < m; v += D::NT) {
            const float4 r = srec[v];
            int lx, ly, lz;
            float dx, dy, dz;
            local_cell<CIC>(r.x, P.off, P.inv_hx, P.box, P.gx_d, P.nx, ox_w, D::NXC, lx, dx);
            local_cell<CIC>(r.y, P.off, P.inv_hy, P.box, P.gy_d, P.ny, y0, D::NYC, ly, dy);
            local_cell<CIC>(r.z, P.off, P.inv_hz, P.box, P.gz_d, P.nz, z0, D::NZC, lz, dz);
            srec[v] = make_float4(dx, dy, dz, r.w);
            if ((unsigned)lx < (unsigned)D::NXC && (unsigned)ly < (unsigned)D::NYC && (unsigned)lz < (unsigned)D::NZC) {
                next16[v] = (uint16_t)atomicExch(&head[(lx * D::NYC + ly) * 32 + lz + 1], (uint32_t)v);
            } else {
                unsigned slot = MAX_OVF;
                if ((unsigned)lx < 16u && (unsigned)ly < 16u && (unsigned)lz < 64u) slot = atomicAdd(&ovf_cnt, 1u);
                if (slot < (unsigned)MAX_OVF) {
                    ovf_v[slot] = (uint16_t)v;
                    ovf_xyz[slot] = (uint16_t)((lx << 10) | (ly << 6) | lz);
                } else {
                    // far outside the tile (arbitrary offset difference): global cell = origin + local
                    deposit_direct(grid, MeshShape{P.nx, P.ny, P.nz, P.x_lo, P.cic}, ldz, slab, abk_wrap_cell(P.x_lo + x0 + lx, P.nx), abk_wrap_cell(y0 + ly, P.ny),
                                   abk_wrap_cell(z0 + lz, P.nz), dx, dy, dz, r.w);
                }
            }
        }
        __syncthreads();
        // ---- queued out-of-domain particles: one warp per particle, one lane per stencil point -----------
        {
            const int novf = (int)min(ovf_cnt, (unsigned)MAX_OVF);
            const int a = lane / 9, b = (lane / 3) % 3, c = lane % 3;
            for (int q = wy; q < novf; q += D::NW) {
                const float4 r = srec[ovf_v[q]];
                const int xyz = ovf_xyz[q];
                const int lx = xyz >> 10, ly = (xyz >> 6) & 15, lz = xyz & 63;
                float wx[3], wyv[3], wz[3];
                mas_w<CIC>(r.x, wx[0], wx[1], wx[2]);
                mas_w<CIC>(r.y, wyv[0], wyv[1], wyv[2]);
                mas_w<CIC>(r.z, wz[0], wz[1], wz[2]);
                if (lane < 27) {
                    const float val = (a == 0 ? wx[0] : (a == 1 ? wx[1] : wx[2])) *
                                      (b == 0 ? wyv[0] : (b == 1 ? wyv[1] : wyv[2])) *
                                      (c == 0 ? wz[0] : (c == 1 ? wz[1] : wz[2])) * r.w;
                    const int64_t gx = slab ? (int64_t)(x0 + lx + a) : (int64_t)abk_wrap_cell(x0 + lx + a - 1, P.nx);
                    const int gy = abk_wrap_cell(y0 + ly + b - 1, P.ny);
                    const int gzo = abk_wrap_cell(z0 + lz + c - 1, P.nz);
                    atomicAdd(grid + gx * sx + (int64_t)gy * ldz + gzo, val);
                }
            }
        }
        // ---- phase 2, lane <-> cell (row wy, lane = local z + 1): rolling 3-plane register window along x --------
        Plane S0, S1, S2;
        plane_zero(S0); plane_zero(S1); plane_zero(S2);
        int to_wrap = slab ? 0x7fffffff : P.nx - gx_first;      // planes left before the periodic wrap (slab grids: none)
        // row y-1 of the plane being emitted; advanced by one plane per step (a running pointer: recomputing the 64-bit
        // address for each of the three reductions of a step cost more instructions than the reductions themselves)
        char *rowp = reinterpret_cast<char *>(grid) + ((int64_t)gx_first * sx + rowo0) * 4;
        const uint32_t *hp = head + wy * 32 + lane;
        // sum the lists of cell column cx into (A, B, C) = planes cx-1, cx, cx+1 (plane index cx, cx+1, cx+2); then
        // plane index cx is complete: combine across lanes, add to the grid, hand the registers back zeroed
        auto step = [&](int cx, Plane &A, Plane &B, Plane &C) {
            if (cx < D::NXC) {
                uint32_t i = hp[cx * (D::NYC * 32)];
                while (i != NIL) {
                    const float4 r = srec[i];
                    const uint16_t nxt = next16[i];
                    i = (nxt == 0xffffu) ? NIL : (uint32_t)nxt;
                    accumulate_particle<CIC>(r, A, B, C);
                }
            }
            const float ctr[3] = {A.q.x, A.s, A.q.y};
            float v[3];
#pragma unroll
            for (int b = 0; b < 3; b++) {
                // neighbour shuffles by one lane: lane 0 (a halo lane, no particles) receives its own zero from below; lane 31
                // receives its own value from above -- zero as well, except EXT, where it owns the cell local z = 30: its
                // upper term belongs to local z = 31 and goes to the stash
                const float up = __shfl_up_sync(0xffffffffu, A.p[b].y, 1);    // lane z-1's contribution to z
                float dn = __shfl_down_sync(0xffffffffu, A.p[b].x, 1);        // lane z+1's contribution to z
                if (EXT && lane == 31) {
                    stash[cx * 3 + b] = A.p[b].y;
                    dn = 0.0f;
                }
                v[b] = ctr[b] + up + dn;
            }
            // one vote per step: rows of zeros (empty tiles of a sparse catalogue) are skipped three at a time; a zero row
            // next to a non-zero one is added as it is (x + 0 = x; the addresses are wrapped into the mesh)
            if (__any_sync(0xffffffffu, (v[0] != 0.0f) | (v[1] != 0.0f) | (v[2] != 0.0f))) {
                atomicAdd(reinterpret_cast<float *>(rowp), v[0]);
                atomicAdd(reinterpret_cast<float *>(rowp + d01b), v[1]);
                atomicAdd(reinterpret_cast<float *>(rowp + d01b + d12b), v[2]);
            }
            plane_zero(A);
            rowp += P.plane_bytes;
            if (--to_wrap == 0) {      // periodic grid: the plane after nx - 1 is plane 0
                to_wrap = P.nx;
                rowp -= P.wrap_bytes;
            }
        };
        {
            int cx = 0;
#pragma unroll 1
            for (; cx + 3 <= D::NPL; cx += 3) {
                step(cx, S0, S1, S2);
                step(cx + 1, S1, S2, S0);
                step(cx + 2, S2, S0, S1);
            }
            if (D::NPL % 3 >= 1) step(cx, S0, S1, S2);
            if (D::NPL % 3 == 2) step(cx + 1, S1, S2, S0);
        }
        // ---- EXT: the upper z-halo cells of this walk (local z = 31): [plane][row] -------------------------------
        if (EXT) {
            __syncwarp();
            const int gzp = abk_wrap_cell(z0 + 31, P.nz);
            for (int e = lane; e < D::STASH; e += 32) {
                const float v = stash[e];
                if (v != 0.0f) {
                    const int b = e % 3, px = e / 3;
                    const int hx = slab ? gx_first + px : abk_wrap_cell(gx_first + px, P.nx);
                    atomicAdd(grid + (int64_t)hx * sx + (int64_t)(b == 0 ? gy0 : (b == 1 ? gy1 : gy2)) * ldz + gzp, v);
                }
            }
            __syncwarp();
        }
    }
}

// validation path: one thread per particle, 27 global reductions in the reference's cell order
__global__ void __launch_bounds__(256) tsc_naive_kernel(const float *__restrict__ pos, const float *__restrict__ w,
                                                        int64_t N, float *__restrict__ grid, TscParams P, int64_t ldz)
{
    const int64_t sx = (int64_t)P.ny * ldz;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        float x = pos[3 * n], y = pos[3 * n + 1], z = pos[3 * n + 2];
        if (P.wrap) { x = wrap_coord(x, P.box); y = wrap_coord(y, P.box); z = wrap_coord(z, P.box); }
        const float W = (w ? w[n] : 1.0f) * P.wscale;
        int cx, cy, cz;
        float dx, dy, dz;
        cells_of(P, x, y, z, cx, cy, cz, dx, dy, dz);
        float wx[3], wy[3], wz[3];
        if (P.cic) {
            cic_w(dx, wx[0], wx[1], wx[2]); cic_w(dy, wy[0], wy[1], wy[2]); cic_w(dz, wz[0], wz[1], wz[2]);
        } else {
            tsc_w(dx, wx[0], wx[1], wx[2]); tsc_w(dy, wy[0], wy[1], wy[2]); tsc_w(dz, wz[0], wz[1], wz[2]);
        }
        for (int a = 0; a < 3; a++) {
            const int64_t gx = abk_wrap_cell(cx + a - 1, P.nx);
            for (int b = 0; b < 3; b++) {
                const int gy = abk_wrap_cell(cy + b - 1, P.ny);
                for (int c = 0; c < 3; c++) {
                    const int gz = abk_wrap_cell(cz + c - 1, P.nz);
                    atomicAdd(grid + gx * sx + (int64_t)gy * ldz + gz, wx[a] * wy[b] * wz[c] * W);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) wrap_inplace_kernel(float *__restrict__ pos, int64_t n3, double box,
                                                           unsigned long long *__restrict__ n_changed)
{
    unsigned long long changed = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = pos[i];
        const float r = wrap_coord(v, box);
        if (r != v && v == v) {
            pos[i] = r;
            changed++;
        }
    }
    if (n_changed && changed) atomicAdd(n_changed, changed);
}

// ---- measurement helper ("micro-benchmarked peak red.global.add.f32 rate ... as the denominator") -----
// Every warp issues `iters` float reductions of 32 lanes into `buf`: mode 0 = coalesced 32-float rows spread over the
// buffer (what the deposit emits), mode 1 = 32 scattered cells per instruction (what one reduction per stencil point,
// the reference's formulation, would be).
__global__ void __launch_bounds__(256) red_rate_kernel(float *__restrict__ buf, int64_t nfloats, int iters, int mode)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nrows = nfloats / 32;
    uint32_t h = (uint32_t)(warp * 32 + lane) * 2654435761u + 12345u;
    for (int it = 0; it < iters; it++) {
        if (mode == 0) {
            const int64_t row = (warp * 7919 + (int64_t)it * nwarps) % nrows;
            atomicAdd(buf + row * 32 + lane, 1.0f);
        } else {
            h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
            atomicAdd(buf + (int64_t)(h % (uint32_t)min(nfloats, (int64_t)0x7fffffff)), 1.0f);
        }
    }
}

// ---- particle routing for an x-slab sharded mesh -------------------------------------------------------
// owner(p) = rank r with xsplit[r] <= cell_x(p) < xsplit[r+1], cell_x = rint(x * f32(nx/box)) mod nx (the
// centre cell of the UNSHIFTED cloud; the half-cell-shifted cloud is handled with one more ghost plane).
// Warp-aggregated counting sort into (x,y,z,w) records grouped by owner.
struct RouteSplit { int v[65]; };

template <bool SCATTER>
__global__ void __launch_bounds__(256) route_kernel(const float *__restrict__ pos, const float *__restrict__ w, int64_t N,
                                                    TscParams P, int nranks, RouteSplit xs,
                                                    unsigned long long *__restrict__ counts, float4 *__restrict__ out)
{
    __shared__ int s_xs[65];
    for (int t = threadIdx.x; t <= nranks; t += blockDim.x) s_xs[t] = xs.v[t];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nround = (N + 31) / 32 * 32;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += (int64_t)gridDim.x * blockDim.x) {
        int owner = -1;
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < N) {
            x = pos[3 * i]; y = pos[3 * i + 1]; z = pos[3 * i + 2];
            if (P.wrap) { x = wrap_coord(x, P.box); y = wrap_coord(y, P.box); z = wrap_coord(z, P.box); }
            int cx, cy, cz;
            float d, d2, d3;
            cells_of(P, x, y, z, cx, cy, cz, d, d2, d3);
            int lo = 0, hi = nranks - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_xs[mid] <= cx) lo = mid; else hi = mid - 1;
            }
            owner = lo;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, owner);
        if (owner < 0) continue;
        const int leader = __ffs(peers) - 1;
        const int rank_in = __popc(peers & ((1u << lane) - 1u));
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(&counts[owner], (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        if (SCATTER) out[base + rank_in] = make_float4(x, y, z, w ? w[i] : 1.0f);
    }
}

// ---- partition_parallel (tsc.py:259-384) ---------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) partition_kernel(const float *__restrict__ pos, const float *__restrict__ w,
                                                        int64_t N, int npart, float inv_pwidth, int coord,
                                                        uint32_t *__restrict__ counts, float *__restrict__ out_pos,
                                                        float *__restrict__ out_w, uint32_t *__restrict__ out_index)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        const float v = coord == 0 ? x : (coord == 1 ? y : z);
        int key = (int)(v * inv_pwidth);  // truncation, like np.int32()
        key = max(0, min(key, npart - 1));
        if (!SCATTER) {
            atomicAdd(&counts[key], 1u);
        } else {
            const uint32_t s = atomicSub(&counts[key], 1u) - 1u;
            out_pos[3 * (int64_t)s] = x;
            out_pos[3 * (int64_t)s + 1] = y;
            out_pos[3 * (int64_t)s + 2] = z;
            if (w) out_w[s] = w[i];
            if (out_index) out_index[s] = (uint32_t)i;
        }
    }
}

__global__ void starts_to_i64_kernel(const uint32_t *__restrict__ excl, int npart, int64_t N, int64_t *__restrict__ starts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npart) starts[i] = excl[i];
    if (i == npart) starts[i] = N;
}

int make_params(const abk_ctx *ctx, TscParams &P, int nx, int ny, int nz, double box, double offset, int wrap, int x_lo, int nxe)
{
    ABK_REQUIRE(nx > 0 && ny > 0 && nz > 0, "grid shape (%d,%d,%d) must be positive", nx, ny, nz);
    ABK_REQUIRE(box > 0, "box must be positive");
    ABK_REQUIRE(nxe > 0 && nxe <= nx + 1 && x_lo >= 0 && x_lo < nx, "bad slab range x_lo=%d nxe=%d nx=%d", x_lo, nxe, nx);
    P.inv_hx = (float)(nx / box);
    P.inv_hy = (float)(ny / box);
    P.inv_hz = (float)(nz / box);
    P.off = (float)offset;
    P.box = box;
    P.nx = nx; P.ny = ny; P.nz = nz;
    P.nxe = nxe; P.x_lo = x_lo;
    const abk_tile_geom g = abk_make_geom(nxe, ny, nz);
    P.nty = g.nty; P.ntz = g.ntz;
    P.wrap = wrap;
    P.cic = ctx->scheme == 1;
    P.gx_d = nx; P.gy_d = ny; P.gz_d = nz;
    P.wscale = ctx->wscale;
    P.plane_bytes = P.wrap_bytes = 0;
    return ABK_OK;
}

int grid_for(const abk_ctx *ctx, int64_t work_items, int threads, int per_sm)
{
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)ctx->num_sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

// ==============================================================================================
extern "C" int abk_wrap_inplace(abk_ctx *ctx, float *pos, int64_t N, double box, int64_t *n_changed_dev)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && (pos || N == 0) && N >= 0, "abk_wrap_inplace: bad arguments");
    if (N == 0) return ABK_OK;
    ABK_LAUNCH(ctx, ABK_K_WRAP, wrap_inplace_kernel<<<grid_for(ctx, 3 * N, 256, 16), 256, 0, ctx->stream>>>(pos, 3 * N, box,
                                                                                 (unsigned long long *)n_changed_dev));
    return ABK_OK;
}

extern "C" int abk_partition_scratch_bytes(int64_t N, int npart, size_t *bytes)
{
    ABK_REQUIRE(bytes && npart > 0 && N >= 0, "abk_partition_scratch_bytes: bad arguments");
    *bytes = abk_align_up((size_t)(npart + 1) * 4, 256) + abk_scan_tmp_bytes(npart);
    return ABK_OK;
}

extern "C" int abk_partition(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int npart, double box,
                             int coord, float *out_pos, float *out_w, int64_t *out_starts, uint32_t *out_index, void *scratch,
                             size_t scratch_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && out_starts && npart > 0 && N >= 0 && coord >= 0 && coord < 3, "abk_partition: bad arguments");
    ABK_REQUIRE(N < (int64_t)1 << 32, "abk_partition: N=%lld exceeds 2^32-1", (long long)N);
    size_t need;
    abk_partition_scratch_bytes(N, npart, &need);
    if (scratch_bytes < need) {
        abk_set_error("abk_partition: scratch %zu < %zu", scratch_bytes, need);
        return ABK_ERR_SCRATCH;
    }
    uint32_t *counts = (uint32_t *)scratch;
    void *tmp = (char *)scratch + abk_align_up((size_t)(npart + 1) * 4, 256);
    ABK_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)(npart + 1) * 4, ctx->stream));
    const float inv_pwidth = (float)((double)npart / box);
    if (N > 0) {
        const int blocks = grid_for(ctx, N, 256, 16);
        ABK_LAUNCH(ctx, ABK_K_PART_HIST, partition_kernel<false><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, npart, inv_pwidth, coord, counts, nullptr, nullptr, nullptr));
        int rc = abk_inclusive_scan_u32(ctx, counts, npart, tmp);
        if (rc) return rc;
        ABK_LAUNCH(ctx, ABK_K_PART_SCATTER, partition_kernel<true><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, npart, inv_pwidth, coord, counts, out_pos, out_w, out_index));
    }
    ABK_LAUNCH(ctx, ABK_K_MISC, starts_to_i64_kernel<<<(npart + 256) / 256, 256, 0, ctx->stream>>>(counts, npart, N, out_starts));
    return ABK_OK;
}

extern "C" int abk_route_particles(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int nx, double box,
                                   int wrap, int nranks, const int32_t *xsplit_h, void *records_out,
                                   int64_t *counts_h)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && (pos || N == 0) && N >= 0 && nranks >= 1 && nranks <= 64 && xsplit_h && counts_h,
                "abk_route_particles: bad arguments");
    ABK_REQUIRE(xsplit_h[0] == 0 && xsplit_h[nranks] == nx, "abk_route_particles: xsplit must run from 0 to nx");
    RouteSplit xs;
    for (int r = 0; r <= nranks; r++) xs.v[r] = xsplit_h[r];
    unsigned long long *counts = ctx->d_scalars + 8;  // 2 x 64 slots would not fit: use 8..39 / 40..
    ABK_REQUIRE(nranks <= 24, "abk_route_particles: at most 24 ranks per node supported");
    unsigned long long *cursors = ctx->d_scalars + 8 + 24;
    ABK_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 48, ctx->stream));
    TscParams P;
    int rcp = make_params(ctx, P, nx, nx, nx, box, 0.0, wrap, 0, nx);
    if (rcp) return rcp;
    unsigned long long h[24];
    if (N > 0) {
        const int blocks = grid_for(ctx, N, 256, 16);
        ABK_LAUNCH(ctx, ABK_K_PART_HIST, route_kernel<false><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, P, nranks, xs, counts, nullptr));
    }
    ABK_CHECK_CUDA(cudaMemcpyAsync(h, counts, sizeof(unsigned long long) * nranks, cudaMemcpyDeviceToHost, ctx->stream));
    ABK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long run = 0, start[24];
    for (int r = 0; r < nranks; r++) { counts_h[r] = (int64_t)h[r]; start[r] = run; run += h[r]; }
    if (N > 0 && records_out) {
        ABK_CHECK_CUDA(cudaMemcpyAsync(cursors, start, sizeof(unsigned long long) * nranks, cudaMemcpyHostToDevice, ctx->stream));
        const int blocks = grid_for(ctx, N, 256, 16);
        ABK_LAUNCH(ctx, ABK_K_PART_SCATTER, route_kernel<true><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, P, nranks, xs, cursors, (float4 *)records_out));
    }
    return ABK_OK;
}

extern "C" int abk_tsc_num_tiles(int nx, int ny, int nz, int64_t *ntiles)
{
    ABK_REQUIRE(ntiles && nx > 0 && ny > 0 && nz > 0, "abk_tsc_num_tiles: bad arguments");
    *ntiles = abk_make_geom(nx, ny, nz).ntiles;
    return ABK_OK;
}

extern "C" int abk_bench_red_rate(abk_ctx *ctx, float *buf, int64_t nfloats, int mode, double *gadds_per_s)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && buf && gadds_per_s && nfloats >= 1024 && (mode == 0 || mode == 1), "abk_bench_red_rate: bad arguments");
    const int blocks = ctx->num_sms * 8, iters = 256;
    cudaEvent_t a, b;
    ABK_CHECK_CUDA(cudaEventCreate(&a));
    ABK_CHECK_CUDA(cudaEventCreate(&b));
    red_rate_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, nfloats, iters, mode);  // warm-up
    ABK_CHECK_CUDA(cudaEventRecord(a, ctx->stream));
    ABK_LAUNCH(ctx, ABK_K_MISC, red_rate_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, nfloats, iters, mode));
    ABK_CHECK_CUDA(cudaEventRecord(b, ctx->stream));
    ABK_CHECK_CUDA(cudaEventSynchronize(b));
    float ms = 0.0f;
    ABK_CHECK_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *gadds_per_s = (double)blocks * 256 * iters / (ms * 1e-3) / 1e9;
    return ABK_OK;
}

extern "C" int abk_tsc_tile_shape(int *tx, int *ty, int *tz)
{
    ABK_REQUIRE(tx && ty && tz, "abk_tsc_tile_shape: null argument");
    *tx = ABK_TX; *ty = ABK_TY; *tz = ABK_TZ;
    return ABK_OK;
}

extern "C" int abk_tsc_bucket_scratch_bytes(int64_t N, int nx, int ny, int nz, size_t *bytes)
{
    ABK_REQUIRE(bytes && nx > 0 && ny > 0 && nz > 0 && N >= 0, "abk_tsc_bucket_scratch_bytes: bad arguments");
    *bytes = abk_scan_tmp_bytes(abk_make_geom(nx, ny, nz).ntiles) + 256;
    return ABK_OK;
}

static int bucket_impl(abk_ctx *ctx, const float *pos, const float *w, int64_t N, const TscParams &P, void *records,
                       uint32_t *tile_starts, void *scratch, size_t scratch_bytes, bool rec4 = false)
{
    const abk_tile_geom g = abk_make_geom(P.nxe, P.ny, P.nz);
    ABK_REQUIRE(g.ntiles < ((int64_t)1 << 31), "too many tiles (%lld)", (long long)g.ntiles);
    ABK_REQUIRE(N <= ((int64_t)1 << 30), "a bucket segment holds at most 2^30 particles (got %lld)", (long long)N);
    size_t need = abk_scan_tmp_bytes(g.ntiles) + 256;
    if (scratch_bytes < need) {
        abk_set_error("abk_tsc_bucket: scratch %zu < %zu", scratch_bytes, need);
        return ABK_ERR_SCRATCH;
    }
    ABK_CHECK_CUDA(cudaMemsetAsync(tile_starts, 0, (size_t)(g.ntiles + 1) * 4, ctx->stream));
    unsigned long long *dropped = ctx->d_scalars + 1;
    if (N > 0) {
        const int vec_ok = (((uintptr_t)pos & 15) == 0) && (!w || ((uintptr_t)w & 15) == 0);
        const int blocks = grid_for(ctx, (N + 3) / 4, 256, 16);
        if (rec4) ABK_LAUNCH(ctx, ABK_K_BUCKET_HIST, tsc_bucket_kernel<false, true><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, P, tile_starts, nullptr, vec_ok, dropped));
        else ABK_LAUNCH(ctx, ABK_K_BUCKET_HIST, tsc_bucket_kernel<false, false><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, P, tile_starts, nullptr, vec_ok, dropped));
        int rc = abk_inclusive_scan_u32(ctx, tile_starts, g.ntiles, scratch);
        if (rc) return rc;
        // sentinel tile_starts[ntiles] = number of bucketed particles = inclusive total
        ABK_CHECK_CUDA(cudaMemcpyAsync(tile_starts + g.ntiles, tile_starts + g.ntiles - 1, 4, cudaMemcpyDeviceToDevice,
                                       ctx->stream));
        if (rec4) ABK_LAUNCH(ctx, ABK_K_BUCKET_SCATTER, tsc_bucket_kernel<true, true><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, P, tile_starts, (float4 *)records, vec_ok, dropped));
        else ABK_LAUNCH(ctx, ABK_K_BUCKET_SCATTER, tsc_bucket_kernel<true, false><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, P, tile_starts, (float4 *)records, vec_ok, dropped));
    }
    return ABK_OK;
}

extern "C" int abk_tsc_bucket(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int nx, int ny, int nz,
                              double box, double offset, int wrap, void *records, uint32_t *tile_starts,
                              void *scratch, size_t scratch_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && records && tile_starts && (pos || N == 0), "abk_tsc_bucket: null argument");
    TscParams P;
    int rc = make_params(ctx, P, nx, ny, nz, box, offset, wrap, 0, nx);
    if (rc) return rc;
    return bucket_impl(ctx, pos, w, N, P, records, tile_starts, scratch, scratch_bytes);
}

extern "C" int abk_tsc_bucket_slab(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int in_records, int nx,
                                   int ny, int nz, double box, double offset, int wrap, int x_lo, int nxe,
                                   void *records, uint32_t *tile_starts, void *scratch, size_t scratch_bytes,
                                   unsigned long long *n_dropped_h)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && records && tile_starts && (pos || N == 0), "abk_tsc_bucket_slab: null argument");
    TscParams P;
    int rc = make_params(ctx, P, nx, ny, nz, box, offset, wrap, x_lo, nxe);
    if (rc) return rc;
    ABK_CHECK_CUDA(cudaMemsetAsync(ctx->d_scalars + 1, 0, 8, ctx->stream));
    rc = bucket_impl(ctx, pos, w, N, P, records, tile_starts, scratch, scratch_bytes, in_records != 0);
    if (rc) return rc;
    if (n_dropped_h) {
        ABK_CHECK_CUDA(cudaMemcpyAsync(n_dropped_h, ctx->d_scalars + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
        ABK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return ABK_OK;
}

typedef void (*deposit_kernel_t)(SegList, float *, TscParams, int64_t, int, int);

extern "C" int abk_tsc_deposit_tiles(abk_ctx *ctx, int nseg, const void *const *records_seg_h,
                                     const uint32_t *const *tile_starts_seg_h, const int64_t *seg_counts_h,
                                     float *grid, int nx, int ny, int nz, int64_t ldz, double box, double offset,
                                     double bucket_offset, int slab, int x_lo, int nxe)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && grid && nseg >= 1 && nseg <= ABK_MAX_SEGMENTS, "abk_tsc_deposit_tiles: bad arguments");
    ABK_REQUIRE(ldz >= nz, "ldz %lld < nz %d", (long long)ldz, nz);
    ABK_REQUIRE(slab || (x_lo == 0 && nxe == nx), "abk_tsc_deposit_tiles: a periodic (non-slab) grid needs x_lo=0, nxe=nx");
    TscParams P;
    int rc = make_params(ctx, P, nx, ny, nz, box, offset, 0, x_lo, nxe);
    if (rc) return rc;
    P.plane_bytes = (int64_t)ny * ldz * 4;
    P.wrap_bytes = (int64_t)nx * P.plane_bytes;
    const abk_tile_geom g = abk_make_geom(nxe, ny, nz);
    ABK_REQUIRE(g.ntx <= 65535 && g.nty <= 65535, "mesh too large for the tile grid (%d x %d tile columns)", g.ntx, g.nty);
    SegList segs;
    segs.nseg = nseg;
    int64_t n_total = 0;
    for (int s = 0; s < nseg; s++) {
        segs.rec[s] = (const float4 *)records_seg_h[s];
        segs.starts[s] = tile_starts_seg_h[s];
        n_total += seg_counts_h ? seg_counts_h[s] : 0;
    }
    // records bucketed at another offset: widen the tile's cell domain by one cell per axis
    const int ext = ((float)bucket_offset != (float)offset) ? 1 : 0;
    const bool cic = ctx->scheme == 1;
    deposit_kernel_t kern = ext ? (cic ? tsc_tile_walk_kernel<1, true> : tsc_tile_walk_kernel<1, false>)
                                : (cic ? tsc_tile_walk_kernel<0, true> : tsc_tile_walk_kernel<0, false>);
    cudaFuncAttributes fa;
    ABK_CHECK_CUDA(cudaFuncGetAttributes(&fa, kern));
    const size_t fixed = walk_smem_bytes(0, ext) + fa.sharedSizeBytes;
    int cap = ctx->tile_capacity & 0xffff;
    if (!cap) {
        // mean occupancy + 5 sigma (Poisson): a uniform catalogue then needs ONE pass per tile; never more than what
        // keeps the target number of CTAs resident (denser tiles take several passes)
        const double mean = g.ntiles > 0 ? (double)n_total / (double)g.ntiles : 0.0;
        cap = (int)((mean + 5.0 * sqrt(mean + 1.0) + 32.0 + 63.0) / 64.0) * 64;
        const int occ = ext ? 3 : 4;
        const int fit = (int)((((size_t)ctx->smem_optin + 1024) / occ - 1024 - fixed) / 18) / 64 * 64;
        if (cap > fit) cap = fit;
        if (cap < 256) cap = 256;
    }
    const size_t smem = walk_smem_bytes(cap, ext);
    ABK_REQUIRE(smem + fa.sharedSizeBytes <= (size_t)ctx->smem_optin, "tile capacity %d needs %zu B shared memory (> %d)", cap,
                smem + fa.sharedSizeBytes, ctx->smem_optin);
    ABK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ABK_LAUNCH(ctx, ABK_K_TILE_DEPOSIT, kern<<<dim3((unsigned)g.ntz, (unsigned)g.nty, (unsigned)g.ntx), ext ? WalkDom<1>::NT : WalkDom<0>::NT, smem, ctx->stream>>>(segs, grid, P, ldz, cap, slab));
    return ABK_OK;
}

extern "C" int abk_tsc_deposit_scratch_bytes(int64_t N, int nx, int ny, int nz, size_t *bytes)
{
    ABK_REQUIRE(bytes && nx > 0 && ny > 0 && nz > 0 && N >= 0, "abk_tsc_deposit_scratch_bytes: bad arguments");
    const int64_t ntiles = abk_make_geom(nx, ny, nz).ntiles;
    const int64_t nseg = (N + ((int64_t)1 << 30) - 1) >> 30;
    size_t b = abk_align_up((size_t)N * 16, 256);                                   // records
    b += (size_t)(nseg > 0 ? nseg : 1) * abk_align_up((size_t)(ntiles + 1) * 4, 256);  // tile_starts per segment
    b += abk_scan_tmp_bytes(ntiles) + 256;
    *bytes = b;
    return ABK_OK;
}

extern "C" int abk_tsc_deposit(abk_ctx *ctx, const float *pos, const float *w, int64_t N, float *grid, int nx, int ny,
                               int nz, int64_t ldz, double box, double offset, int wrap, void *scratch,
                               size_t scratch_bytes)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && grid && (pos || N == 0) && N >= 0, "abk_tsc_deposit: bad arguments");
    if (N == 0) return ABK_OK;
    size_t need;
    int rc = abk_tsc_deposit_scratch_bytes(N, nx, ny, nz, &need);
    if (rc) return rc;
    if (scratch_bytes < need) {
        abk_set_error("abk_tsc_deposit: scratch %zu < %zu", scratch_bytes, need);
        return ABK_ERR_SCRATCH;
    }
    const int64_t ntiles = abk_make_geom(nx, ny, nz).ntiles;
    const int64_t SEG = (int64_t)1 << 30;
    const int nseg = (int)((N + SEG - 1) / SEG);
    ABK_REQUIRE(nseg <= ABK_MAX_SEGMENTS, "N=%lld needs more than %d segments", (long long)N, ABK_MAX_SEGMENTS);
    char *p = (char *)scratch;
    float4 *records = (float4 *)p;
    p += abk_align_up((size_t)N * 16, 256);
    const void *rec_h[ABK_MAX_SEGMENTS];
    const uint32_t *starts_h[ABK_MAX_SEGMENTS];
    int64_t cnt_h[ABK_MAX_SEGMENTS];
    uint32_t *starts0 = (uint32_t *)p;
    p += (size_t)nseg * abk_align_up((size_t)(ntiles + 1) * 4, 256);
    void *scan_tmp = p;
    const size_t scan_bytes = abk_scan_tmp_bytes(ntiles) + 256;
    for (int s = 0; s < nseg; s++) {
        const int64_t b = (int64_t)s * SEG, n = (N - b < SEG) ? N - b : SEG;
        uint32_t *starts = (uint32_t *)((char *)starts0 + (size_t)s * abk_align_up((size_t)(ntiles + 1) * 4, 256));
        rc = abk_tsc_bucket(ctx, pos + 3 * b, w ? w + b : nullptr, n, nx, ny, nz, box, offset, wrap, records + b, starts,
                            scan_tmp, scan_bytes);
        if (rc) return rc;
        rec_h[s] = records + b;
        starts_h[s] = starts;
        cnt_h[s] = n;
    }
    return abk_tsc_deposit_tiles(ctx, nseg, rec_h, starts_h, cnt_h, grid, nx, ny, nz, ldz, box, offset, offset, 0, 0, nx);
}

extern "C" int abk_tsc_deposit_naive(abk_ctx *ctx, const float *pos, const float *w, int64_t N, float *grid, int nx,
                                     int ny, int nz, int64_t ldz, double box, double offset, int wrap)
{
    abk_device_guard entry_guard(ctx ? ctx->device : -1);
    ABK_REQUIRE(ctx && grid && (pos || N == 0) && N >= 0, "abk_tsc_deposit_naive: bad arguments");
    if (N == 0) return ABK_OK;
    TscParams P;
    int rc = make_params(ctx, P, nx, ny, nz, box, offset, wrap, 0, nx);
    if (rc) return rc;
    ABK_LAUNCH(ctx, ABK_K_NAIVE_DEPOSIT, tsc_naive_kernel<<<grid_for(ctx, N, 256, 16), 256, 0, ctx->stream>>>(pos, w, N, grid, P, ldz));
    return ABK_OK;
}
