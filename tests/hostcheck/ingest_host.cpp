// Host build of the per-record decoders in abacusutils_b200/csrc/abk_ingest.cuh (test infrastructure only).
// Compiled by tests/test_ingest_host.py with g++ -O2 -ffp-contract=off so that the exact rounding sequence the
// CUDA kernels execute can be compared bit for bit with the unmodified reference on the CPU.
#include <cmath>
#include <cstdint>
#include <limits>

#include "abk_ingest.cuh"

template <typename T>
static void rvint(const int32_t *in, int64_t n, double box, T *pos, T *vel)
{
    const double posscale = box / 1e6;
    for (int64_t i = 0; i < 3 * n; i++) {
        if (pos) pos[i] = abk_rvint_pos<T>(in[i], posscale);
        if (vel) vel[i] = abk_rvint_vel<T>(in[i]);
    }
}

template <typename T>
static int64_t pack9(const uint8_t *data, int64_t nrec, double box, double velz, T *pos, T *vel)
{
    abk_pack9_cell<T> h;
    const T nan = std::numeric_limits<T>::quiet_NaN();
    h.pscale = h.cellx = h.celly = h.cellz = h.vscale = nan;
    int64_t w = 0;
    for (int64_t i = 0; i < nrec; i++) {
        int s[6];
        abk_pack9_expand(data + 9 * i, s);
        if (abk_pack9_is_header(data + 9 * i)) {
            h = abk_pack9_header<T>(s, (T)box, (T)velz);
        } else {
            T p[3], v[3];
            abk_pack9_particle<T>(s, h, p, v);
            for (int a = 0; a < 3; a++) {
                if (pos) pos[3 * w + a] = p[a];
                if (vel) vel[3 * w + a] = v[a];
            }
            w++;
        }
    }
    return w;
}

extern "C" {
void hc_rvint_f32(const int32_t *in, int64_t n, double box, float *pos, float *vel) { rvint<float>(in, n, box, pos, vel); }
void hc_rvint_f64(const int32_t *in, int64_t n, double box, double *pos, double *vel) { rvint<double>(in, n, box, pos, vel); }
int64_t hc_pack9_f32(const uint8_t *d, int64_t n, double box, double velz, float *pos, float *vel) { return pack9<float>(d, n, box, velz, pos, vel); }
int64_t hc_pack9_f64(const uint8_t *d, int64_t n, double box, double velz, double *pos, double *vel) { return pack9<double>(d, n, box, velz, pos, vel); }
}

template <typename T>
static void pids(const uint64_t *packed, int64_t n, double box, int64_t ppd, int64_t *pid, T *lagr_pos, int16_t *lagr_idx,
                 uint8_t *tagged, T *density)
{
    const T inv_ppd = (T)(box / (double)ppd), half = (T)(box / 2);
    for (int64_t i = 0; i < n; i++) {
        if (pid) pid[i] = abk_pid_pid(packed[i]);
        if (lagr_pos) abk_pid_lagr_pos<T>(packed[i], inv_ppd, half, lagr_pos + 3 * i);
        if (lagr_idx) abk_pid_lagr_idx(packed[i], lagr_idx + 3 * i);
        if (tagged) tagged[i] = abk_pid_tagged(packed[i]);
        if (density) density[i] = abk_pid_density<T>(packed[i]);
    }
}

extern "C" {
void hc_pids_f32(const uint64_t *p, int64_t n, double box, int64_t ppd, int64_t *pid, float *lp, int16_t *li, uint8_t *tg, float *de) { pids<float>(p, n, box, ppd, pid, lp, li, tg, de); }
void hc_pids_f64(const uint64_t *p, int64_t n, double box, int64_t ppd, int64_t *pid, double *lp, int16_t *li, uint8_t *tg, double *de) { pids<double>(p, n, box, ppd, pid, lp, li, tg, de); }
}
