// Per-record arithmetic of the Abacus particle decoders, shared by the CUDA kernels in
// abk_ingest.cu and by the host-compiled check in tests/ingest_host.cpp (g++ -ffp-contract=off), which pins
// every rounding step against the unmodified reference on the CPU:
//   RVint   abacusnbody/data/bitpacked.py:101-120   (_unpack_rvint)
//   pack9   abacusnbody/data/pack9.py:58-123        (_unpack_pack9, _expand_to_short)
// The reference kernels are Numba without fastmath: no FMA contraction, so products and sums are rounded
// separately (abk_mul / abk_add below).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ABK_HD __host__ __device__ __forceinline__
#else
#define ABK_HD inline
#endif

ABK_HD float abk_mul(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
ABK_HD float abk_add(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
ABK_HD double abk_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
ABK_HD double abk_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// ---- RVint: three int32 per particle, 20 bits of position (signed, units of box/1e6) over 12 bits of
// velocity (offset 2048, units of 6000/2048 km/s).  The reference promotes int32 op uint32 to int64 and
// multiplies by a float64 scale; the store casts to the output dtype (bitpacked.py:110-120).
template <typename T>
ABK_HD T abk_rvint_pos(int32_t v, double posscale)
{
    return (T)((double)((int64_t)v >> 12) * posscale);
}
template <typename T>
ABK_HD T abk_rvint_vel(int32_t v)
{
    return (T)((double)(((int64_t)v & 0xFFF) - 2048) * (6000.0 / 2048));
}

// ---- pack9: 9 bytes = six 12-bit fields, each stored with an offset of 2048 (pack9.py:110-123)
ABK_HD void abk_pack9_expand(const uint8_t *c, int s[6])
{
    s[0] = ((c[1] & 0x0F) | (c[0] << 4)) - 2048;
    s[1] = (((c[1] & 0xF0) << 4) | c[2]) - 2048;
    s[2] = ((c[4] & 0x0F) | (c[3] << 4)) - 2048;
    s[3] = (((c[4] & 0xF0) << 4) | c[5]) - 2048;
    s[4] = ((c[7] & 0x0F) | (c[6] << 4)) - 2048;
    s[5] = (((c[7] & 0xF0) << 4) | c[8]) - 2048;
}

ABK_HD bool abk_pack9_is_header(const uint8_t *c) { return c[0] == 0xFF; }

// cell header -> the five numbers a particle record needs (pack9.py:84-93), in the output dtype T
template <typename T>
struct abk_pack9_cell {
    T pscale, cellx, celly, cellz, vscale;
};

template <typename T>
ABK_HD abk_pack9_cell<T> abk_pack9_header(const int s[6], T boxsize, T velz)
{
    abk_pack9_cell<T> h;
    const T invcpd = (T)(1.0 / (double)(s[1] + 2000));
    const T csize = abk_mul(boxsize, invcpd);
    const double halfbox = (double)boxsize / 2;   // exact in either precision
    h.vscale = abk_mul(abk_mul((T)((double)(s[2] + 2000) * 0.0005), invcpd), velz);
    h.cellx = (T)abk_add(abk_mul((double)s[3] + 2000.5, (double)csize), -halfbox);
    h.celly = (T)abk_add(abk_mul((double)s[4] + 2000.5, (double)csize), -halfbox);
    h.cellz = (T)abk_add(abk_mul((double)s[5] + 2000.5, (double)csize), -halfbox);
    h.pscale = (T)(0.0005 * (double)csize);
    return h;
}

// particle record: int16 field times a T scale, plus the cell origin, both rounded in T (pack9.py:99-105)
template <typename T>
ABK_HD void abk_pack9_particle(const int s[6], const abk_pack9_cell<T> &h, T pos[3], T vel[3])
{
    pos[0] = abk_add(abk_mul((T)s[0], h.pscale), h.cellx);
    pos[1] = abk_add(abk_mul((T)s[1], h.pscale), h.celly);
    pos[2] = abk_add(abk_mul((T)s[2], h.pscale), h.cellz);
    vel[0] = abk_mul((T)s[3], h.vscale);
    vel[1] = abk_mul((T)s[4], h.vscale);
    vel[2] = abk_mul((T)s[5], h.vscale);
}

// ---- packed PID / aux word (bitpacked.py:15-23, 269-311): bits 0-14, 16-30, 32-46 = Lagrangian index (x, y, z),
// bit 48 = L2-tagged, bits 49-58 = sqrt of the local density.  uint64 times a float scale promotes to float64 in the
// reference; the store rounds once to the output type.
constexpr uint64_t ABK_AUX_X = 0x7FFFull, ABK_AUX_Y = 0x7FFF0000ull, ABK_AUX_Z = 0x7FFF00000000ull;
constexpr uint64_t ABK_AUX_PID = ABK_AUX_X | ABK_AUX_Y | ABK_AUX_Z, ABK_AUX_DENS = 0x07FE000000000000ull;

ABK_HD void abk_pid_lagr_idx(uint64_t p, int16_t idx[3])
{
    idx[0] = (int16_t)(p & ABK_AUX_X);
    idx[1] = (int16_t)((p & ABK_AUX_Y) >> 16);
    idx[2] = (int16_t)((p & ABK_AUX_Z) >> 32);
}
// inv_ppd and half are already rounded to T by the caller (bitpacked.py:286-287)
template <typename T>
ABK_HD void abk_pid_lagr_pos(uint64_t p, T inv_ppd, T half, T pos[3])
{
    pos[0] = (T)abk_add(abk_mul((double)(p & ABK_AUX_X), (double)inv_ppd), -(double)half);
    pos[1] = (T)abk_add(abk_mul((double)((p & ABK_AUX_Y) >> 16), (double)inv_ppd), -(double)half);
    pos[2] = (T)abk_add(abk_mul((double)((p & ABK_AUX_Z) >> 32), (double)inv_ppd), -(double)half);
}
ABK_HD uint8_t abk_pid_tagged(uint64_t p) { return (uint8_t)((p >> 48) & 1ull); }
template <typename T>
ABK_HD T abk_pid_density(uint64_t p)
{
    const uint64_t r = (p & ABK_AUX_DENS) >> 49;
    return (T)(r * r);
}
ABK_HD int64_t abk_pid_pid(uint64_t p) { return (int64_t)(p & ABK_AUX_PID); }
