// TSC mass assignment on B200 (sm_100a).
//
// Replaces, with a different algorithm, the reference's Numba kernels
//   _wrap_inplace        analysis/tsc.py:219-226
//   partition_parallel   analysis/tsc.py:259-384
//   _tsc_parallel        analysis/tsc.py:229-256   (two-colour x-stripe schedule)
//   _tsc_scatter         analysis/tsc.py:394-507   (27 read-modify-writes per particle)
//
// Design:
//   A. bucket particles by the (8 x 8 x 32)-cell tile of their cloud's centre cell: histogram
//      (one 4-byte reduction per particle), scan, scatter of 16-byte (x,y,z,w) records.  The open
//      write frontier is one 128-byte line per tile; it stays (mostly) resident in the 126 MB L2.
//   B. one CTA per tile: per-cell particle lists are built in shared memory with one integer
//      exchange per particle; then warp = y-row, lane = z-cell, x walked serially: every lane sums
//      the 27 stencil weights of ITS cell's particles in registers (a rolling 3-plane window along
//      x), neighbouring lanes are combined with two shuffles, neighbouring rows through a
//      shared-memory tile written without atomics (each (plane,row) is owned by exactly one warp
//      per phase).  Shared-memory float atomics are a CAS loop on sm_100a (ATOMS.CAST.SPIN), so the
//      kernel uses none.  The finished tile + 1-cell halo is added to the grid with coalesced float
//      reductions (REDG.ADD.F32): ~1.7 per CELL instead of 27 per PARTICLE.
//   C. interlacing: the half-cell-shifted deposit reuses the records bucketed for the unshifted grid
//      (template EXT: one more cell row/plane per tile; z overflow is queued and deposited
//      warp-cooperatively).  CIC (analysis/cic.py) is the same update with other weights (template CIC).
// Kernel variants kept for measurement (abk_ctx_set_tile_capacity bits 16-18): precomputed 32-byte
// records (PRE), private per-warp slabs (PRIV), row streaming without an output tile; the default
// (16-byte records, shared tile) is the fastest on B200 (profiles/r1_ncu_summary.md).
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
    int flush_v2;                  // experiment: flush tile rows with 8-byte vector reductions (red.global.add.v2.f32)
    float wscale;                  // multiplies every weight as the bucket records are written (1 unless the caller
                                   // folds the field normalisation into the deposit, abk_ctx_set_weight_scale)
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
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (ok[q]) records[slot[q]] = make_float4(c[3 * q], c[3 * q + 1], c[3 * q + 2], wv[q] * P.wscale);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tile deposit.  Geometry of the shared-memory working set of one CTA (= one tile of 8 x 8 x 32 cells):
//   slab[w]   : warp w (= y-row w of the tile) owns a PRIVATE float slab [TX+2 planes][3 rows][TZ+2],
//               the image of its row's clouds on rows w-1, w, w+1 (+1-cell halo in x and z).  Private
//               slabs make the main loop free of block barriers and of atomics.
//   head[c]   : first particle of cell c's list (index into srec) or NIL
//   srec[2v]  : (wx-, wx0, wx+, dz)            precomputed in the lane<->particle load pass, so the
//   srec[2v+1]: (wy- W, wy0 W, wy+ W, next)     divergent lane<->cell loop is 2 LDS.128 + 43 FP ops
// EXT = 1 widens the cell domain of a tile by one cell in x and y (and treats z-overflow with direct
// reductions): that is what the half-cell-shifted (interlaced) deposit needs when it REUSES the records
// bucketed for the unshifted grid -- the shifted centre cell is the unshifted one or its +1 neighbour.
template <int EXT>
struct TileDom {
    static constexpr int NXC = ABK_TX + EXT, NYC = ABK_TY + EXT;       // cells with particle lists
    static constexpr int OX = NXC + 2, OY = NYC + 2, OZ = ABK_TZ + 2;  // output region incl. 1-cell halo
    static constexpr int OUT_N = OX * OY * OZ;
    static constexpr int SLAB_PLANE = 3 * OZ, SLAB_WORDS = OX * SLAB_PLANE;
    static constexpr int NCELL = NXC * NYC * ABK_TZ;
    static constexpr int NW = NYC, NT = NW * 32;  // one warp per y-row
};
constexpr uint32_t NIL = 0xffffffffu;

struct SegList {
    int nseg;
    const float4 *rec[ABK_MAX_SEGMENTS];
    const uint32_t *starts[ABK_MAX_SEGMENTS];
};

// PRE : 32-byte records with precomputed x/y weights (fewer instructions in the divergent loop)
//       vs 16-byte (dx,dy,dz,W) records + u16 links (less shared memory -> more resident CTAs)
// PRIV: private per-warp slabs (no barriers in the main loop) vs one shared output tile with a
//       block barrier after every row phase
static size_t deposit_smem_bytes(int cap, bool pre, bool priv, int ext)
{
    const size_t out_n = ext ? TileDom<1>::OUT_N : TileDom<0>::OUT_N;
    const size_t slab = ext ? (size_t)TileDom<1>::NW * TileDom<1>::SLAB_WORDS : (size_t)TileDom<0>::NW * TileDom<0>::SLAB_WORDS;
    const size_t ncell = ext ? TileDom<1>::NCELL : TileDom<0>::NCELL;
    const size_t out = (priv ? slab : out_n) * 4;
    const size_t rec = pre ? (size_t)cap * 32 : (size_t)cap * 16 + (size_t)cap * 2;
    return abk_align_up(out + ncell * 4, 16) + rec + 64;
}

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

// two adjacent floats in one 8-byte reduction (sm_90+); p must be 8-byte aligned
__device__ __forceinline__ void red_add_v2(float *p, float a, float b)
{
#if defined(__CUDA_ARCH__)
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
#else
    atomicAdd(p, a);
    atomicAdd(p + 1, b);
#endif
}

// Add one finished x-plane of this lane's register window to the output.
// S[b][c]: contribution of cell (row w, z = lane) to row w-1+b, cell z-1+c.  Lane z receives the
// c=+1 term of lane z-1 and the c=-1 term of lane z+1; the two halo cells are lanes 0 / 31's.
// PRIV: `dst` is the warp's private slab [plane][3 rows][OZ] and the first pass stores instead of adding.
// shared: `dst` is the CTA's tile [plane][OY][OZ]; rows of different warps are distinct within a
//         phase (b) but not across phases, hence the block barrier after each phase.
template <typename D, bool PRIV>
__device__ __forceinline__ void emit_plane(float *__restrict__ dst, const float (&S)[3][3], int x, int wy, int lane,
                                           bool first)
{
    float *plane = dst + (x + 1) * (PRIV ? D::SLAB_PLANE : D::OY * D::OZ);
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const float up = __shfl_up_sync(0xffffffffu, S[b][2], 1);
        const float dn = __shfl_down_sync(0xffffffffu, S[b][0], 1);
        const float v = S[b][1] + (lane > 0 ? up : 0.0f) + (lane < 31 ? dn : 0.0f);
        float *row = plane + (PRIV ? b : wy + b) * D::OZ;
        if (PRIV && first) {  // the first pass writes every slab entry exactly once: no zero-fill, no RMW
            row[lane + 1] = v;
            if (lane == 0) row[0] = S[b][0];
            if (lane == 31) row[D::OZ - 1] = S[b][2];
        } else {
            row[lane + 1] += v;
            if (lane == 0) row[0] += S[b][0];
            if (lane == 31) row[D::OZ - 1] += S[b][2];
        }
        if (!PRIV) __syncthreads();
    }
}

// 27 global reductions for a particle whose centre cell lies outside the cell domain of the tile it
// was bucketed in (only possible when the deposit offset differs from the bucketing offset).
__device__ __noinline__ void deposit_direct(float *__restrict__ grid, const TscParams &P, int64_t ldz, int slab,
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

// MINB (0 = by variant): minimum resident CTAs per SM the register allocation must allow.  The default kernel is built
// twice: for 4 CTAs/SM (64 registers, 60 bytes of spills) and for 3 (80 registers, no spills); the launch picks the
// second whenever shared memory limits the SM to three CTAs anyway (e.g. ~1 particle per cell, one pass per tile).
template <bool PRE, bool PRIV, int EXT, bool CIC, int MINB = 0>
__global__ void __launch_bounds__(TileDom<EXT>::NT, MINB ? MINB : (PRE ? 2 : ((PRIV || EXT) ? 3 : 4)))
tsc_tile_deposit_kernel(SegList segs, float *__restrict__ grid, TscParams P, int64_t ldz, int cap, int slab)
{
    using D = TileDom<EXT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int OUT_WORDS = PRIV ? D::NW * D::SLAB_WORDS : D::OUT_N;
    constexpr int HEAD_OFF = OUT_WORDS;
    constexpr int REC_OFF_BYTES = ((OUT_WORDS + D::NCELL) * 4 + 15) / 16 * 16;
    float *outbuf = reinterpret_cast<float *>(smem_raw);
    uint32_t *head = reinterpret_cast<uint32_t *>(outbuf + HEAD_OFF);
    float4 *srec = reinterpret_cast<float4 *>(smem_raw + REC_OFF_BYTES);
    uint16_t *next16 = reinterpret_cast<uint16_t *>(srec + cap);  // !PRE only
    __shared__ uint32_t seg_beg[ABK_MAX_SEGMENTS], seg_cnt[ABK_MAX_SEGMENTS], seg_off[ABK_MAX_SEGMENTS + 1];
    __shared__ const float4 *seg_rec[ABK_MAX_SEGMENTS];
    // particles whose cell falls just outside the tile's cell domain (z overflow of the shifted deposit):
    // queued here and deposited warp-cooperatively, one lane per stencil point
    constexpr int MAX_OVF = 512;
    __shared__ uint16_t ovf_v[MAX_OVF], ovf_xyz[MAX_OVF];
    __shared__ unsigned ovf_cnt;

    const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
    const uint32_t tile = blockIdx.x;
    if (tid < segs.nseg) {
        const uint32_t b = segs.starts[tid][tile], e = segs.starts[tid][tile + 1];
        seg_beg[tid] = b;
        seg_cnt[tid] = e - b;
        seg_rec[tid] = segs.rec[tid];
    }
    __syncthreads();
    uint32_t total = 0;
    for (int s = 0; s < segs.nseg; s++) total += seg_cnt[s];
    if (total == 0) return;
    if (tid == 0) {
        uint32_t run = 0;
        for (int s = 0; s < segs.nseg; s++) { seg_off[s] = run; run += seg_cnt[s]; }
        seg_off[segs.nseg] = run;
    }

    const uint32_t tz = tile % P.ntz, ty = (tile / P.ntz) % P.nty, tx = tile / (P.ntz * P.nty);
    const int x0 = tx * ABK_TX, y0 = ty * ABK_TY, z0 = tz * ABK_TZ;  // x0 relative to x_lo

    float *myslab = outbuf + wy * D::SLAB_WORDS;  // PRIV only
    if (!PRIV)
        for (int i = tid; i < D::OUT_N; i += D::NT) outbuf[i] = 0.0f;

    for (uint32_t chunk0 = 0; chunk0 < total; chunk0 += cap) {
        const int m = (int)min((uint32_t)cap, total - chunk0);
        const bool first = (chunk0 == 0);
        for (int c = tid; c < D::NCELL; c += D::NT) head[c] = NIL;
        if (tid == 0) ovf_cnt = 0;
        __syncthreads();
        // ---- lane <-> particle: per-cell lists (and, PRE, the x/y weights once per particle) -------
        const int ox_g = P.x_lo + x0;  // global cell index of the tile origin in x (slab: may exceed nx, handled by wrap)
        {
            int sg = 0;  // this thread's virtual indices only grow: the segment lookup is incremental
            for (int v0 = tid; v0 < m; v0 += 4 * D::NT) {
                float4 rr[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {  // issue the (streaming) record loads of four particles first
                    const int vq = v0 + q * D::NT;
                    if (vq < m) {
                        const uint32_t u = chunk0 + vq;
                        while (u >= seg_off[sg + 1]) sg++;
                        rr[q] = __ldcs(seg_rec[sg] + seg_beg[sg] + (u - seg_off[sg]));
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int v = v0 + q * D::NT;
                    if (v >= m) break;
                    const float4 r = rr[q];
                    int lx, ly, lz;
                    float dx, dy, dz;
                    local_cell<CIC>(r.x, P.off, P.inv_hx, P.box, P.gx_d, P.nx, ox_g >= P.nx ? ox_g - P.nx : ox_g, D::NXC, lx, dx);
                    local_cell<CIC>(r.y, P.off, P.inv_hy, P.box, P.gy_d, P.ny, y0, D::NYC, ly, dy);
                    local_cell<CIC>(r.z, P.off, P.inv_hz, P.box, P.gz_d, P.nz, z0, ABK_TZ, lz, dz);
                    if ((unsigned)lx < (unsigned)D::NXC && (unsigned)ly < (unsigned)D::NYC && (unsigned)lz < (unsigned)ABK_TZ) {
                        const int c = (lx * D::NYC + ly) * ABK_TZ + lz;
                        if (PRE) {
                            float wxm, wx0, wxp, wym, wy0, wyp;
                            mas_w<CIC>(dx, wxm, wx0, wxp);
                            mas_w<CIC>(dy, wym, wy0, wyp);
                            srec[2 * v] = make_float4(wxm, wx0, wxp, dz);
                            const uint32_t old = atomicExch(&head[c], (uint32_t)v);
                            srec[2 * v + 1] = make_float4(wym * r.w, wy0 * r.w, wyp * r.w, __uint_as_float(old));
                        } else {
                            srec[v] = make_float4(dx, dy, dz, r.w);
                            next16[v] = (uint16_t)atomicExch(&head[c], (uint32_t)v);
                        }
                    } else {
                        unsigned slot = MAX_OVF;
                        if ((unsigned)lx < 16u && (unsigned)ly < 16u && (unsigned)lz < 64u) slot = atomicAdd(&ovf_cnt, 1u);
                        if (slot < (unsigned)MAX_OVF) {
                            srec[PRE ? 2 * v : v] = make_float4(dx, dy, dz, r.w);
                            ovf_v[slot] = (uint16_t)v;
                            ovf_xyz[slot] = (uint16_t)((lx << 10) | (ly << 6) | lz);
                        } else {
                            // far outside the tile (arbitrary offset difference): global cell = origin + local
                            deposit_direct(grid, P, ldz, slab, abk_wrap_cell(P.x_lo + x0 + lx, P.nx), abk_wrap_cell(y0 + ly, P.ny),
                                           abk_wrap_cell(z0 + lz, P.nz), dx, dy, dz, r.w);
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- queued out-of-domain particles: one warp per particle, one lane per stencil point -----------
        {
            const int novf = (int)min(ovf_cnt, (unsigned)MAX_OVF);
            const int64_t sxo = (int64_t)P.ny * ldz;
            const int a = lane / 9, b = (lane / 3) % 3, c = lane % 3;
            for (int q = wy; q < novf; q += D::NW) {
                const float4 r = srec[PRE ? 2 * ovf_v[q] : ovf_v[q]];
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
                    const int gz = abk_wrap_cell(z0 + lz + c - 1, P.nz);
                    atomicAdd(grid + gx * sxo + (int64_t)gy * ldz + gz, val);
                }
            }
        }
        // ---- lane <-> cell (row wy, z = lane); rolling 3-plane register window along x ---------------
        float S0[3][3], S1[3][3], S2[3][3];
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int c = 0; c < 3; c++) S0[b][c] = S1[b][c] = S2[b][c] = 0.0f;

#pragma unroll 1
        for (int cx = 0; cx < D::NXC; cx++) {
            uint32_t i = head[(cx * D::NYC + wy) * ABK_TZ + lane];
            while (i != NIL) {
                float wx[3], wyW[3], wz[3];
                if (PRE) {
                    const float4 A = srec[2 * i], B = srec[2 * i + 1];
                    i = __float_as_uint(B.w);
                    wx[0] = A.x; wx[1] = A.y; wx[2] = A.z;
                    wyW[0] = B.x; wyW[1] = B.y; wyW[2] = B.z;
                    mas_w<CIC>(A.w, wz[0], wz[1], wz[2]);
                } else {
                    const float4 r = srec[i];
                    const uint16_t nxt = next16[i];
                    i = (nxt == 0xffffu) ? NIL : (uint32_t)nxt;
                    mas_w<CIC>(r.x, wx[0], wx[1], wx[2]);
                    mas_w<CIC>(r.y, wyW[0], wyW[1], wyW[2]);
                    mas_w<CIC>(r.z, wz[0], wz[1], wz[2]);
                    wyW[0] *= r.w; wyW[1] *= r.w; wyW[2] *= r.w;
                }
#pragma unroll
                for (int b = 0; b < 3; b++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float t = wyW[b] * wz[c];
                        S0[b][c] = fmaf(wx[0], t, S0[b][c]);
                        S1[b][c] = fmaf(wx[1], t, S1[b][c]);
                        S2[b][c] = fmaf(wx[2], t, S2[b][c]);
                    }
            }
            emit_plane<D, PRIV>(PRIV ? myslab : outbuf, S0, cx - 1, wy, lane, first);
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int c = 0; c < 3; c++) { S0[b][c] = S1[b][c]; S1[b][c] = S2[b][c]; S2[b][c] = 0.0f; }
        }
        emit_plane<D, PRIV>(PRIV ? myslab : outbuf, S0, D::NXC - 1, wy, lane, first);
        emit_plane<D, PRIV>(PRIV ? myslab : outbuf, S1, D::NXC, wy, lane, first);
        if (PRIV) __syncthreads();  // lists are rebuilt by the next pass; slabs are complete for the flush
    }

    // ---- flush tile + halo with float reductions -----------------------------------------------------
    // Warp w owns output rows oy = w, w + NW, ... of every x-plane; lanes on z, so the reductions of
    // one instruction hit 32 consecutive floats.  Everything that does not depend on the plane (row
    // pointers of the <= 3 contributing slabs, wrapped y/z indices) is hoisted.
    const int64_t sx = (int64_t)P.ny * ldz;
    // vector flush needs the whole 32-cell row inside the mesh (no z wrap inside it) and 8-byte aligned pairs
    const bool pair_ok = !PRIV && P.flush_v2 && z0 + ABK_TZ <= P.nz && (ldz & 1) == 0 && (((uintptr_t)grid) & 7) == 0;
    const int gz = abk_wrap_cell(z0 + lane, P.nz);
    const int ozh = lane ? D::OZ - 1 : 0;
    const int gzh = abk_wrap_cell(z0 + ozh - 1, P.nz);
    for (int oy = wy; oy < D::OY; oy += D::NW) {
        const int gy = abk_wrap_cell(y0 + oy - 1, P.ny);
        // PRIV: contributing (warp, row) pairs: w = oy - b for b = 0..2 with 0 <= w < NW
        const bool ok0 = oy < D::NW, ok1 = oy >= 1 && oy - 1 < D::NW, ok2 = oy >= 2 && oy - 2 < D::NW;
        const float *p0 = outbuf + (ok0 ? oy : 0) * D::SLAB_WORDS;
        const float *p1 = outbuf + (ok1 ? oy - 1 : 0) * D::SLAB_WORDS + D::OZ;
        const float *p2 = outbuf + (ok2 ? oy - 2 : 0) * D::SLAB_WORDS + 2 * D::OZ;
        int gx = slab ? x0 : abk_wrap_cell(x0 - 1, P.nx);
        for (int ox = 0; ox < D::OX; ox++) {
            float v = 0.0f, vh = 0.0f;
            if (PRIV) {
                const int o = ox * D::SLAB_PLANE;
                if (ok0) { v += p0[o + lane + 1]; if (lane < 2) vh += p0[o + ozh]; }
                if (ok1) { v += p1[o + lane + 1]; if (lane < 2) vh += p1[o + ozh]; }
                if (ok2) { v += p2[o + lane + 1]; if (lane < 2) vh += p2[o + ozh]; }
            } else {
                const float *r = outbuf + (ox * D::OY + oy) * D::OZ;
                v = r[lane + 1];
                if (lane < 2) vh = r[ozh];
            }
            float *dst = grid + gx * sx + (int64_t)gy * ldz;
            if (pair_ok) {
                // lanes 0..15 take the 32 interior cells two at a time, lanes 16/17 the two z-halo cells
                const float *r = outbuf + (ox * D::OY + oy) * D::OZ;
                if (lane < 16) {
                    const float a = r[2 * lane + 1], b = r[2 * lane + 2];
                    if (a != 0.0f || b != 0.0f) red_add_v2(dst + z0 + 2 * lane, a, b);
                } else if (lane < 18) {
                    const int oz = (lane == 16) ? 0 : D::OZ - 1;
                    const float h = r[oz];
                    if (h != 0.0f) atomicAdd(dst + abk_wrap_cell(z0 + oz - 1, P.nz), h);
                }
            } else {
                if (v != 0.0f) atomicAdd(dst + gz, v);
                if (lane < 2 && vh != 0.0f) atomicAdd(dst + gzh, vh);
            }
            gx++;
            if (!slab && gx >= P.nx) gx -= P.nx;
        }
    }
}

// ==============================================================================================
// Tile deposit, round-2 formulation ("walk" kernel; the default).
//   phase 1 (lane <-> particle): records of the tile are loaded once, converted to (dx, dy, dz, W) + local cell, and
//            threaded onto per-cell lists with one shared-memory exchange each (ATOMS.EXCH runs at ~1.3e12 lane-ops/s on
//            B200, scripts/micro/deposit_micro.cu -- it is not the cost centre round 1 took it for);
//   phase 2 (lane <-> cell): warp = y-row of the tile, lane = z-cell, x walked serially.  Every lane sums the clouds of
//            the particles of ITS cell into a rolling window of three x-planes held in registers, stored as packed
//            float pairs so the 27 multiply-adds of a particle are 12 FFMA2 + 3 FFMA with scalar-broadcast operands
//            (36 scalar FP instructions in round 1).  A finished plane is combined with the neighbouring lanes by two
//            shuffles per row and added STRAIGHT to the grid: three coalesced 32-float reductions per x-step.  There is
//            no shared-memory output tile and no block barrier inside the walk: the warps of a CTA run independently, so
//            an x-step costs the longest of the warp's 32 lists, not of the CTA's 256.  The two z-halo cells of every
//            row go through a 60-float per-warp stash and are reduced at the end of the walk (the reduction rate is per
//            instruction, ~7 SM-cycles each whether 2 or 32 lanes are active).
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
    float2 X2, Y2, Z2;
    float x0, y0, z0;
    mas_pair<CIC>(r.x, 1.0f, X2, x0);
    mas_pair<CIC>(r.y, r.w, Y2, y0);
    mas_pair<CIC>(r.z, 1.0f, Z2, z0);
    float2 T[3];
    T[0] = __fmul2_rn(Z2, make_float2(Y2.x, Y2.x));
    T[1] = __fmul2_rn(Z2, make_float2(y0, y0));
    T[2] = __fmul2_rn(Z2, make_float2(Y2.y, Y2.y));
    const float2 U = __fmul2_rn(Y2, make_float2(z0, z0));
    const float u0 = y0 * z0;
    plane_fma(A, X2.x, T, U, u0);
    plane_fma(B, x0, T, U, u0);
    plane_fma(C, X2.y, T, U, u0);
}

template <int EXT>
struct WalkDom {
    static constexpr int NXC = ABK_TX + EXT, NYC = ABK_TY + EXT;  // cells with particle lists (x, y); z: one per lane
    static constexpr int NCELL = NXC * NYC * 32;
    static constexpr int NW = NYC, NT = NW * 32;                  // one warp per y-row
    static constexpr int NPL = NXC + 2;                           // x-planes a row walk emits (1-cell halo either side)
    static constexpr int STASH = NPL * 3 * 2;                     // z-halo values of one walk: [plane][row][side]
};

static size_t walk_smem_bytes(int cap, int ext)
{
    const size_t ncell = ext ? WalkDom<1>::NCELL : WalkDom<0>::NCELL;
    const size_t stash = ext ? (size_t)WalkDom<1>::NW * WalkDom<1>::STASH : (size_t)WalkDom<0>::NW * WalkDom<0>::STASH;
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
    // particles whose cell falls just outside the tile's cell domain (z overflow of the shifted deposit):
    // queued here and deposited warp-cooperatively, one lane per stencil point
    constexpr int MAX_OVF = 256;
    __shared__ uint16_t ovf_v[MAX_OVF], ovf_xyz[MAX_OVF];
    __shared__ unsigned ovf_cnt;

    const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
    const uint32_t tile = blockIdx.x;
    if (tid == 0) {
        uint32_t run = 0;
        for (int s = 0; s < segs.nseg; s++) {
            const uint32_t b = segs.starts[s][tile], e = segs.starts[s][tile + 1];
            seg_beg[s] = b;
            seg_off[s] = run;
            run += e - b;
        }
        seg_off[segs.nseg] = run;
#if defined(__CUDA_ARCH__)
        mbar_init(&bar, 1);
#endif
    }
    __syncthreads();
    const uint32_t total = seg_off[segs.nseg];
    if (total == 0) return;

    const uint32_t tz = tile % P.ntz, ty = (tile / P.ntz) % P.nty, tx = tile / (P.ntz * P.nty);
    const int x0 = tx * ABK_TX, y0 = ty * ABK_TY, z0 = tz * ABK_TZ;  // x0 relative to x_lo
    const int64_t sx = (int64_t)P.ny * ldz;

    // plane-independent pieces of this warp's output addresses: rows y0+wy-1 .. y0+wy+1, column z0+lane.  Rows or columns
    // beyond the mesh (ragged last tile, meshes smaller than a tile) only ever carry zeros, which are not written.
    const int gz = abk_wrap_cell(z0 + lane, P.nz);
    int rowo[3];
#pragma unroll
    for (int b = 0; b < 3; b++) rowo[b] = (int)((int64_t)abk_wrap_cell(y0 + wy + b - 1, P.ny) * ldz) + gz;
    float *stash = stash_all + wy * D::STASH;
    const int gx_first = slab ? x0 : wrap_near(x0 - 1, P.nx);  // global x of plane 0 (local x = -1)
    uint32_t parity = 0;

    for (uint32_t chunk0 = 0; chunk0 < total; chunk0 += cap) {
        const int m = (int)min((uint32_t)cap, total - chunk0);
        if (chunk0) __syncthreads();  // the previous pass' walks are done with the lists and records
        // ---- stage the raw records of this pass: one bulk copy per segment that overlaps [chunk0, chunk0 + m) ----
#if defined(__CUDA_ARCH__)
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bar, (uint32_t)m * 16u);
            for (int s = 0; s < segs.nseg; s++) {
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
#pragma unroll 2
        for (int v = tid; v < m; v += D::NT) {
            const float4 r = srec[v];
            int lx, ly, lz;
            float dx, dy, dz;
            local_cell<CIC>(r.x, P.off, P.inv_hx, P.box, P.gx_d, P.nx, ox_w, D::NXC, lx, dx);
            local_cell<CIC>(r.y, P.off, P.inv_hy, P.box, P.gy_d, P.ny, y0, D::NYC, ly, dy);
            local_cell<CIC>(r.z, P.off, P.inv_hz, P.box, P.gz_d, P.nz, z0, 32, lz, dz);
            srec[v] = make_float4(dx, dy, dz, r.w);
            if ((unsigned)lx < (unsigned)D::NXC && (unsigned)ly < (unsigned)D::NYC && (unsigned)lz < 32u) {
                next16[v] = (uint16_t)atomicExch(&head[(lx * D::NYC + ly) * 32 + lz], (uint32_t)v);
            } else {
                unsigned slot = MAX_OVF;
                if ((unsigned)lx < 16u && (unsigned)ly < 16u && (unsigned)lz < 64u) slot = atomicAdd(&ovf_cnt, 1u);
                if (slot < (unsigned)MAX_OVF) {
                    ovf_v[slot] = (uint16_t)v;
                    ovf_xyz[slot] = (uint16_t)((lx << 10) | (ly << 6) | lz);
                } else {
                    // far outside the tile (arbitrary offset difference): global cell = origin + local
                    deposit_direct(grid, P, ldz, slab, abk_wrap_cell(P.x_lo + x0 + lx, P.nx), abk_wrap_cell(y0 + ly, P.ny),
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
        // ---- phase 2, lane <-> cell (row wy, z = lane): rolling 3-plane register window along x --------------
        Plane S0, S1, S2;
        plane_zero(S0); plane_zero(S1); plane_zero(S2);
        int gx = gx_first;
        float *pl = grid + (int64_t)gx * sx;
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
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const float up = __shfl_up_sync(0xffffffffu, A.p[b].y, 1);    // lane z-1's contribution to z
                const float dn = __shfl_down_sync(0xffffffffu, A.p[b].x, 1);  // lane z+1's contribution to z
                const float v = ctr[b] + (lane > 0 ? up : 0.0f) + (lane < 31 ? dn : 0.0f);
                if (v != 0.0f) atomicAdd(pl + rowo[b], v);
                if (lane == 0) stash[(cx * 3 + b) * 2] = A.p[b].x;        // z = -1
                if (lane == 31) stash[(cx * 3 + b) * 2 + 1] = A.p[b].y;   // z = 32
            }
            plane_zero(A);
            gx++;
            pl += sx;
            if (!slab && gx >= P.nx) { gx -= P.nx; pl -= (int64_t)P.nx * sx; }
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
        // ---- the z-halo cells of this walk: [plane][row][side] -----------------------------------------------
        __syncwarp();
        {
            const int gzm = abk_wrap_cell(z0 - 1, P.nz), gzp = abk_wrap_cell(z0 + 32, P.nz);
            for (int e = lane; e < D::STASH; e += 32) {
                const float v = stash[e];
                if (v != 0.0f) {
                    const int side = e & 1, b = (e >> 1) % 3, px = (e >> 1) / 3;
                    const int hx = slab ? gx_first + px : abk_wrap_cell(gx_first + px, P.nx);
                    atomicAdd(grid + (int64_t)hx * sx + (rowo[b == 0 ? 0 : (b == 1 ? 1 : 2)] - gz) + (side ? gzp : gzm), v);
                }
            }
        }
        __syncwarp();
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
                                                        float *__restrict__ out_w)
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
    P.flush_v2 = ctx->flush_v2;
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
                             int coord, float *out_pos, float *out_w, int64_t *out_starts, void *scratch,
                             size_t scratch_bytes)
{
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
        ABK_LAUNCH(ctx, ABK_K_PART_HIST, partition_kernel<false><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, npart, inv_pwidth, coord, counts, nullptr, nullptr));
        int rc = abk_inclusive_scan_u32(ctx, counts, npart, tmp);
        if (rc) return rc;
        ABK_LAUNCH(ctx, ABK_K_PART_SCATTER, partition_kernel<true><<<blocks, 256, 0, ctx->stream>>>(pos, w, N, npart, inv_pwidth, coord, counts, out_pos, out_w));
    }
    ABK_LAUNCH(ctx, ABK_K_MISC, starts_to_i64_kernel<<<(npart + 256) / 256, 256, 0, ctx->stream>>>(counts, npart, N, out_starts));
    return ABK_OK;
}

extern "C" int abk_route_particles(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int nx, double box,
                                   int wrap, int nranks, const int32_t *xsplit_h, void *records_out,
                                   int64_t *counts_h)
{
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

// ==============================================================================================
// Two-level bucketing (experiment, abk_tsc_bucket2): same result layout as abk_tsc_bucket (records grouped by
// tile, tile_starts exclusive), but the 16-byte records reach memory in coalesced runs instead of one scattered
// store per particle -- the classic scatter is bound by the rate of scattered store requests, not by bandwidth.
//   level 1: tile id >> shift  (<= 1024 coarse buckets, each a contiguous range of tiles)
//   level 2: tile id within the coarse bucket (<= 1024 tiles)
// Each level is a CTA-local multisplit of a 4096-record chunk: rank inside (chunk, bucket) with one shared-memory
// atomic per record, one global cursor reservation per (chunk, non-empty bucket), records staged in shared memory
// in bucket order and written out as runs of consecutive addresses.
namespace {

constexpr int SPLIT_CH = 4096, SPLIT_NT = 256, SPLIT_IT = SPLIT_CH / SPLIT_NT, SPLIT_MAXB = 1024;

struct SplitSmem {
    float4 stage[SPLIT_CH];
    uint16_t skey[SPLIT_CH];
    uint32_t cnt[SPLIT_MAXB], off[SPLIT_MAXB], gbase[SPLIT_MAXB];
    uint32_t warp_tot[SPLIT_NT / 32];
};

// One chunk: `rec[it]`/`key[it]` are this thread's items (item it of thread t is chunk element it*256 + t, valid if
// below m).  cursors[key] hands out positions in `out`.
__device__ __forceinline__ void multisplit_chunk(SplitSmem &S, const float4 (&rec)[SPLIT_IT], const int (&key)[SPLIT_IT], int m,
                                                 int nb, uint32_t *__restrict__ cursors, float4 *__restrict__ out)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int b = tid; b < nb; b += SPLIT_NT) S.cnt[b] = 0;
    __syncthreads();
    uint32_t rank[SPLIT_IT];
#pragma unroll
    for (int it = 0; it < SPLIT_IT; it++)
        if (it * SPLIT_NT + tid < m) rank[it] = atomicAdd(&S.cnt[key[it]], 1u);
    __syncthreads();
    // exclusive scan of cnt[0..nb) (4 buckets per thread), and the global reservation of every non-empty bucket
    uint32_t c[4], run = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int b = tid * 4 + q;
        c[q] = b < nb ? S.cnt[b] : 0u;
        run += c[q];
    }
    uint32_t inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) S.warp_tot[wid] = inc;
    __syncthreads();
    uint32_t base = inc - run;
    for (int w = 0; w < wid; w++) base += S.warp_tot[w];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int b = tid * 4 + q;
        if (b < nb) {
            S.off[b] = base;
            if (c[q]) S.gbase[b] = atomicAdd(&cursors[b], c[q]);
            base += c[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SPLIT_IT; it++)
        if (it * SPLIT_NT + tid < m) {
            const uint32_t slot = S.off[key[it]] + rank[it];
            S.stage[slot] = rec[it];
            S.skey[slot] = (uint16_t)key[it];
        }
    __syncthreads();
    for (int i = tid; i < m; i += SPLIT_NT) {
        const int b = S.skey[i];
        out[S.gbase[b] + ((uint32_t)i - S.off[b])] = S.stage[i];
    }
    __syncthreads();
}

// level 1: particles (pos/w) -> temp records grouped by coarse bucket
__global__ void __launch_bounds__(SPLIT_NT, 2)
split_coarse_kernel(const float *__restrict__ pos, const float *__restrict__ w, int64_t N, TscParams P, int shift, int nb,
                    uint32_t *__restrict__ cursors, float4 *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char split_raw[];
    SplitSmem &S = *reinterpret_cast<SplitSmem *>(split_raw);
    const int64_t nchunks = (N + SPLIT_CH - 1) / SPLIT_CH;
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int64_t base = ch * SPLIT_CH;
        const int m = (int)min((int64_t)SPLIT_CH, N - base);
        float4 rec[SPLIT_IT];
        int key[SPLIT_IT];
#pragma unroll
        for (int it = 0; it < SPLIT_IT; it++) {
            const int e = it * SPLIT_NT + threadIdx.x;
            key[it] = 0;
            if (e < m) {
                const int64_t i = base + e;
                float x = __ldcs(pos + 3 * i), y = __ldcs(pos + 3 * i + 1), z = __ldcs(pos + 3 * i + 2);
                if (P.wrap) { x = wrap_coord(x, P.box); y = wrap_coord(y, P.box); z = wrap_coord(z, P.box); }
                uint32_t tile = 0;
                tile_of(P, x, y, z, tile);
                rec[it] = make_float4(x, y, z, (w ? __ldcs(w + i) : 1.0f) * P.wscale);
                key[it] = (int)(tile >> shift);
            }
        }
        multisplit_chunk(S, rec, key, m, nb, cursors, out);
    }
}

// level 2: every coarse bucket (a contiguous record range of `tmp`) is refined by `splits` CTAs; keys are tile ids
// relative to the bucket's first tile, cursors are the global per-tile cursors
__global__ void __launch_bounds__(SPLIT_NT, 2)
split_fine_kernel(const float4 *__restrict__ tmp, const uint32_t *__restrict__ tile_starts, int64_t ntiles, TscParams P, int shift,
                  int splits, uint32_t *__restrict__ cursors, float4 *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char split_raw[];
    SplitSmem &S = *reinterpret_cast<SplitSmem *>(split_raw);
    const int cb = blockIdx.x / splits, part = blockIdx.x % splits;
    const int64_t t0 = (int64_t)cb << shift;
    const int64_t t1 = min(t0 + ((int64_t)1 << shift), ntiles);
    const int nb = (int)(t1 - t0);
    const int64_t r0 = tile_starts[t0], r1 = tile_starts[t1];
    const int64_t nchunks = (r1 - r0 + SPLIT_CH - 1) / SPLIT_CH;
    for (int64_t ch = part; ch < nchunks; ch += splits) {
        const int64_t base = r0 + ch * SPLIT_CH;
        const int m = (int)min((int64_t)SPLIT_CH, r1 - base);
        float4 rec[SPLIT_IT];
        int key[SPLIT_IT];
#pragma unroll
        for (int it = 0; it < SPLIT_IT; it++) {
            const int e = it * SPLIT_NT + threadIdx.x;
            key[it] = 0;
            if (e < m) {
                rec[it] = __ldcs(tmp + base + e);
                uint32_t tile = 0;
                tile_of(P, rec[it].x, rec[it].y, rec[it].z, tile);
                key[it] = (int)((int64_t)tile - t0);
            }
        }
        multisplit_chunk(S, rec, key, m, nb, cursors + t0, out);
    }
}

// incl[t] (inclusive scan of the tile histogram) -> starts[t] exclusive with starts[ntiles] = total, a second copy as the
// per-tile cursors, and the coarse cursors coarse[c] = starts[c << shift]
__global__ void __launch_bounds__(256) split_starts_kernel(const uint32_t *__restrict__ incl, int64_t ntiles, int shift,
                                                           uint32_t *__restrict__ starts, uint32_t *__restrict__ cur_fine,
                                                           uint32_t *__restrict__ cur_coarse)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = t ? incl[t - 1] : 0u;
        starts[t] = v;
        if (t < ntiles) {
            cur_fine[t] = v;
            if ((t & (((int64_t)1 << shift) - 1)) == 0) cur_coarse[t >> shift] = v;
        }
    }
}

int split_shift(int64_t ntiles)
{
    int bits = 0;
    while (((int64_t)1 << bits) < ntiles) bits++;
    return (bits + 1) / 2;
}

}  // namespace

extern "C" int abk_tsc_bucket2_scratch_bytes(int64_t N, int nx, int ny, int nz, size_t *bytes)
{
    ABK_REQUIRE(bytes && nx > 0 && ny > 0 && nz > 0 && N >= 0, "abk_tsc_bucket2_scratch_bytes: bad arguments");
    const int64_t ntiles = abk_make_geom(nx, ny, nz).ntiles;
    *bytes = abk_align_up(abk_scan_tmp_bytes(ntiles) + 256, 256) + 2 * abk_align_up((size_t)(ntiles + 1) * 4, 256) +
             abk_align_up((size_t)SPLIT_MAXB * 4, 256) + abk_align_up((size_t)(N > 0 ? N : 1) * 16, 256);
    return ABK_OK;
}

extern "C" int abk_tsc_bucket2(abk_ctx *ctx, const float *pos, const float *w, int64_t N, int nx, int ny, int nz, double box,
                               double offset, int wrap, void *records, uint32_t *tile_starts, void *scratch,
                               size_t scratch_bytes)
{
    ABK_REQUIRE(ctx && records && tile_starts && (pos || N == 0), "abk_tsc_bucket2: null argument");
    TscParams P;
    int rc = make_params(ctx, P, nx, ny, nz, box, offset, wrap, 0, nx);
    if (rc) return rc;
    const abk_tile_geom g = abk_make_geom(nx, ny, nz);
    const int shift = split_shift(g.ntiles);
    const int64_t ncoarse = (g.ntiles + ((int64_t)1 << shift) - 1) >> shift;
    // meshes with more than 2^20 tiles would need a third level: use the one-level scatter there
    if (N == 0 || ((int64_t)1 << shift) > SPLIT_MAXB || ncoarse > SPLIT_MAXB)
        return bucket_impl(ctx, pos, w, N, P, records, tile_starts, scratch, scratch_bytes);
    ABK_REQUIRE(N <= ((int64_t)1 << 30), "a bucket segment holds at most 2^30 particles (got %lld)", (long long)N);
    size_t need = 0;
    abk_tsc_bucket2_scratch_bytes(N, nx, ny, nz, &need);
    if (scratch_bytes < need || ((uintptr_t)scratch & 255)) {
        abk_set_error("abk_tsc_bucket2: scratch %zu < %zu or not 256-byte aligned", scratch_bytes, need);
        return ABK_ERR_SCRATCH;
    }
    char *sp = (char *)scratch;
    void *scan_tmp = sp;                   sp += abk_align_up(abk_scan_tmp_bytes(g.ntiles) + 256, 256);
    uint32_t *incl = (uint32_t *)sp;       sp += abk_align_up((size_t)(g.ntiles + 1) * 4, 256);
    uint32_t *cur_fine = (uint32_t *)sp;   sp += abk_align_up((size_t)(g.ntiles + 1) * 4, 256);
    uint32_t *cur_coarse = (uint32_t *)sp; sp += abk_align_up((size_t)SPLIT_MAXB * 4, 256);
    float4 *tmp = (float4 *)sp;

    ABK_CHECK_CUDA(cudaMemsetAsync(incl, 0, (size_t)(g.ntiles + 1) * 4, ctx->stream));
    const int vec_ok = (((uintptr_t)pos & 15) == 0) && (!w || ((uintptr_t)w & 15) == 0);
    const int hblocks = grid_for(ctx, (N + 3) / 4, 256, 16);
    ABK_LAUNCH(ctx, ABK_K_BUCKET_HIST, tsc_bucket_kernel<false, false><<<hblocks, 256, 0, ctx->stream>>>(
                                           pos, w, N, P, incl, nullptr, vec_ok, ctx->d_scalars + 1));
    rc = abk_inclusive_scan_u32(ctx, incl, g.ntiles, scan_tmp);
    if (rc) return rc;
    ABK_LAUNCH(ctx, ABK_K_MISC, split_starts_kernel<<<grid_for(ctx, g.ntiles + 1, 256, 8), 256, 0, ctx->stream>>>(
                                    incl, g.ntiles, shift, tile_starts, cur_fine, cur_coarse));
    const size_t smem = sizeof(SplitSmem);
    ABK_CHECK_CUDA(cudaFuncSetAttribute(split_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ABK_CHECK_CUDA(cudaFuncSetAttribute(split_fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nchunks = (N + SPLIT_CH - 1) / SPLIT_CH;
    const int64_t cap = (int64_t)ctx->num_sms * 2;
    ABK_LAUNCH(ctx, ABK_K_BUCKET_SCATTER, split_coarse_kernel<<<(unsigned)(nchunks < cap ? nchunks : cap), SPLIT_NT, smem, ctx->stream>>>(
                                              pos, w, N, P, shift, (int)ncoarse, cur_coarse, tmp));
    int splits = (int)((cap * 4 + ncoarse - 1) / ncoarse);
    if (splits < 1) splits = 1;
    ABK_LAUNCH(ctx, ABK_K_BUCKET_SCATTER, split_fine_kernel<<<(unsigned)(ncoarse * splits), SPLIT_NT, smem, ctx->stream>>>(
                                              tmp, tile_starts, g.ntiles, P, shift, splits, cur_fine, (float4 *)records));
    return ABK_OK;
}

// kernel variant (PRE, PRIV): abk_ctx_set_tile_capacity's bits 16..18 select it for experiments

static int pick_capacity(const abk_ctx *ctx, int64_t n_total, int64_t ntiles, bool pre, bool priv, int ext, int per_sm)
{
    if (ctx->tile_capacity & 0xffff) return ctx->tile_capacity & 0xffff;
    // mean occupancy + 5 sigma (Poisson): a uniform catalogue then needs ONE pass per tile.  A second pass
    // repeats the whole per-cell work, which costs more than one resident CTA less per SM (measured on
    // B200, config 3: 64 ms with 3 CTAs/SM and one pass vs 68 ms with 4 CTAs/SM and two passes for 40%
    // of the tiles), so the capacity may take one CTA/SM; denser tiles simply take several passes.
    const double mean = ntiles > 0 ? (double)n_total / (double)ntiles : 0.0;
    const double want = mean + 5.0 * sqrt(mean + 1.0) + 32.0;
    int cap = (int)((want + 63.0) / 64.0) * 64;
    const size_t fixed = deposit_smem_bytes(0, pre, priv, ext);
    int fit = 256;
    for (int occ = per_sm; occ >= (per_sm > 2 ? per_sm - 1 : per_sm); occ--) {
        const size_t per_cta = (size_t)(ctx->smem_optin + 1024) / occ - 1024;
        fit = (int)((per_cta - fixed) / (pre ? 32 : 18)) / 64 * 64;
        if (cap <= fit) break;
    }
    if (cap > fit) cap = fit;
    if (cap < 256) cap = 256;
    return cap;
}

typedef void (*deposit_kernel_t)(SegList, float *, TscParams, int64_t, int, int);

static deposit_kernel_t pick_kernel(bool, bool, int ext, bool cic)
{
    if (cic) return ext ? tsc_tile_deposit_kernel<false, false, 1, true> : tsc_tile_deposit_kernel<false, false, 0, true>;
    return ext ? tsc_tile_deposit_kernel<false, false, 1, false> : tsc_tile_deposit_kernel<false, false, 0, false>;
}

extern "C" int abk_tsc_deposit_tiles(abk_ctx *ctx, int nseg, const void *const *records_seg_h,
                                     const uint32_t *const *tile_starts_seg_h, const int64_t *seg_counts_h,
                                     float *grid, int nx, int ny, int nz, int64_t ldz, double box, double offset,
                                     double bucket_offset, int slab, int x_lo, int nxe)
{
    ABK_REQUIRE(ctx && grid && nseg >= 1 && nseg <= ABK_MAX_SEGMENTS, "abk_tsc_deposit_tiles: bad arguments");
    ABK_REQUIRE(ldz >= nz, "ldz %lld < nz %d", (long long)ldz, nz);
    ABK_REQUIRE(slab || (x_lo == 0 && nxe == nx), "abk_tsc_deposit_tiles: a periodic (non-slab) grid needs x_lo=0, nxe=nx");
    TscParams P;
    int rc = make_params(ctx, P, nx, ny, nz, box, offset, 0, x_lo, nxe);
    if (rc) return rc;
    const abk_tile_geom g = abk_make_geom(nxe, ny, nz);
    SegList segs;
    segs.nseg = nseg;
    int64_t n_total = 0;
    for (int s = 0; s < nseg; s++) {
        segs.rec[s] = (const float4 *)records_seg_h[s];
        segs.starts[s] = tile_starts_seg_h[s];
        n_total += seg_counts_h ? seg_counts_h[s] : 0;
    }
    // records bucketed at another offset: widen the tile's cell domain by one cell in x and y
    const int ext = ((float)bucket_offset != (float)offset) ? 1 : 0;
    const bool cic = ctx->scheme == 1;
    const int variant = (ctx->tile_capacity >> 16) & 7;  // 0 = walk kernel (default), 1 = round-1 shared-tile kernel (A/B)
    if (variant == 0) {
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
        ABK_LAUNCH(ctx, ABK_K_TILE_DEPOSIT, kern<<<(unsigned)g.ntiles, ext ? WalkDom<1>::NT : WalkDom<0>::NT, smem, ctx->stream>>>(segs, grid, P, ldz, cap, slab));
        return ABK_OK;
    }
    const bool pre = false, priv = false;
    const int per_sm = ext ? 3 : 4;
    const int cap = pick_capacity(ctx, n_total, g.ntiles, pre, priv, ext, per_sm);
    const size_t smem = deposit_smem_bytes(cap, pre, priv, ext);
    ABK_REQUIRE((int)smem <= ctx->smem_optin, "tile capacity %d needs %zu B shared memory (> %d)", cap, smem, ctx->smem_optin);
    deposit_kernel_t kern = pick_kernel(pre, priv, ext, cic);
    // resident CTAs per SM allowed by shared memory (1 KB per CTA is reserved by the driver)
    const int occ_smem = (int)((size_t)(ctx->smem_optin + 1024) / (smem + 1024));
    if (!ext && !cic && occ_smem <= 3 && !ctx->no_minb3) kern = tsc_tile_deposit_kernel<false, false, 0, false, 3>;
    const int threads = ext ? TileDom<1>::NT : TileDom<0>::NT;
    ABK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ABK_LAUNCH(ctx, ABK_K_TILE_DEPOSIT, kern<<<(unsigned)g.ntiles, threads, smem, ctx->stream>>>(segs, grid, P, ldz, cap, slab));
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
    ABK_REQUIRE(ctx && grid && (pos || N == 0) && N >= 0, "abk_tsc_deposit_naive: bad arguments");
    if (N == 0) return ABK_OK;
    TscParams P;
    int rc = make_params(ctx, P, nx, ny, nz, box, offset, wrap, 0, nx);
    if (rc) return rc;
    ABK_LAUNCH(ctx, ABK_K_NAIVE_DEPOSIT, tsc_naive_kernel<<<grid_for(ctx, N, 256, 16), 256, 0, ctx->stream>>>(pos, w, N, grid, P, ldz));
    return ABK_OK;
}
