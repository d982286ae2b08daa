// Micro-benchmark 6 (round 2): two-pass multisplit with PERSISTENT shared-memory staging.
// scatter_micro5 showed that the one-pass bucket scatter pays a DRAM fill for every half-written 32-byte sector.  Here a
// CTA keeps a small queue per bucket in shared memory across batches and only ever flushes whole, 32-byte-aligned sectors
// (pairs of 16-byte records, one st.global.v8.f32 per pair); one global cursor reservation per flush instead of one per record.
// Two levels (coarse = tile / D, then the tile inside the coarse bucket), each <= 1024 buckets.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/split_micro scripts/micro/split_micro.cu && /tmp/split_micro
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int TX = 8, TY = 8, TZ = 30;
struct Geo { int n, nty, ntz; uint32_t ntiles; float fn; };

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__device__ __forceinline__ uint32_t tile_of(const Geo &g, float x, float y, float z)
{
    int cx = (int)rintf(x * g.fn), cy = (int)rintf(y * g.fn), cz = (int)rintf(z * g.fn);
    cx = cx >= g.n ? cx - g.n : cx; cy = cy >= g.n ? cy - g.n : cy; cz = cz >= g.n ? cz - g.n : cz;
    return ((uint32_t)(cx / TX) * g.nty + (uint32_t)(cy / TY)) * g.ntz + (uint32_t)(cz / TZ);
}

__global__ void fill_pos(float *pos, int64_t n3)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x)
        pos[i] = (hash32((uint32_t)i * 2654435761u + (uint32_t)(i >> 32)) >> 8) * (1.0f / 16777216.0f);
}

__global__ void hist_kernel(const float *__restrict__ pos, int64_t N, Geo g, uint32_t *__restrict__ counts)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < N; q += (int64_t)gridDim.x * blockDim.x) {
        const float4 *p4 = reinterpret_cast<const float4 *>(pos + 12 * q);
        const float4 a = __ldcs(p4), b = __ldcs(p4 + 1), d = __ldcs(p4 + 2);
        const float c[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
#pragma unroll
        for (int k = 0; k < 4; k++) atomicAdd(&counts[tile_of(g, c[3 * k], c[3 * k + 1], c[3 * k + 2])], 1u);
    }
}

// exclusive starts -> cursors (rounded up to an even record index: the odd leading slot of a bucket is filled last)
__global__ void cursors_kernel(const uint32_t *__restrict__ starts, uint32_t ntiles, int D, uint32_t *__restrict__ cur_fine, uint32_t *__restrict__ cur_coarse)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) {
        const uint32_t s = starts[t];
        cur_fine[t] = s + (s & 1u);
        if (t % D == 0) cur_coarse[t / D] = s + (s & 1u);
    }
}

__device__ __forceinline__ void st_pair(float4 *p, const float4 a, const float4 b)
{
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y),
                 "f"(b.z), "f"(b.w)
                 : "memory");
}

// Writes the first n staged records of bucket b (slot-major staging: record i of bucket b at stage[i * pitch + b]) at the
// positions base .. base+n-1 of the bucket's region [S, E); a position == E means the odd leading slot S.
template <int MODE>
__device__ __forceinline__ void flush_bucket(const float4 *__restrict__ stage, int pitch, int b, uint32_t n, uint32_t base, uint32_t S, uint32_t E,
                                             float4 *__restrict__ out)
{
    if (MODE == 2) { base &= 0xfffeu; S = 0; E = 0x7fffffffu; }
    if (MODE >= 4) { S = 0; E = 0x7fffffffu; }      // experiment: all stores inside a 1 MB window
    for (uint32_t i = 0; i < n; i += 2) {
        const float4 r0 = stage[i * pitch + b];
        const uint32_t idx = base + i;
        if (i + 1 < n) {
            const float4 r1 = stage[(i + 1) * pitch + b];
            if (MODE == 1) { if (r0.x == -5.f && r1.x == -7.f) out[0] = r0; }      // experiment: no stores
            else if (!(idx & 1u) && idx + 1 < E) st_pair(out + idx, r0, r1);
            else {
                out[idx >= E ? S : idx] = r0;
                out[idx + 1 >= E ? S : idx + 1] = r1;
            }
        } else out[idx >= E ? S : idx] = r0;
    }
}

template <int LEVEL, int NT, int R, int FMIN, int MODE>
__global__ void __launch_bounds__(NT, NT == 1024 ? 1 : 2)
split_kernel(const float *__restrict__ pos, const float4 *__restrict__ tmp, int64_t N, Geo g, const uint32_t *__restrict__ starts,
             uint32_t *__restrict__ cursors, float4 *__restrict__ out, int D, int ncoarse, int parts, int slots)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    int nb, t0 = 0;
    int64_t lo = 0, hi = 0;
    if (LEVEL == 0) {
        nb = ncoarse;
    } else {
        const int cb = blockIdx.x / parts, part = blockIdx.x % parts;
        t0 = cb * D;
        const int t1 = min(t0 + D, (int)g.ntiles);
        nb = t1 - t0;
        const int64_t r0 = starts[t0], r1 = starts[t1];
        const int64_t per = (((r1 - r0) + parts - 1) / parts + 1) & ~(int64_t)1;
        lo = min(r0 + part * per, r1);
        hi = min(lo + per, r1);
    }
    const int pitch = (nb + 7) & ~7;
    const int C = min(32, slots / pitch);
    float4 *stage = reinterpret_cast<float4 *>(smem_raw);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(stage + slots);
    uint32_t *bS = cnt + 1024, *bE = bS + 1024;
    for (int b = tid; b < nb; b += NT) {
        cnt[b] = 0;
        bE[1024 + b] = 0;
        if (LEVEL == 0) {
            bS[b] = starts[min((int64_t)b * D, (int64_t)g.ntiles)];
            bE[b] = starts[min((int64_t)(b + 1) * D, (int64_t)g.ntiles)];
        } else {
            bS[b] = starts[t0 + b];
            bE[b] = starts[t0 + b + 1];
        }
    }
    __syncthreads();

    constexpr int BATCH = NT * R;
    constexpr int NRAW = LEVEL == 0 ? (3 * R) / 4 : R;
    int64_t nbatch, bi, bstep;
    if (LEVEL == 0) { nbatch = (N + BATCH - 1) / BATCH; bi = blockIdx.x; bstep = gridDim.x; }
    else { nbatch = (hi - lo + BATCH - 1) / BATCH; bi = 0; bstep = 1; }

    float4 raw[NRAW];
    auto load = [&](int64_t b_) {
        if (LEVEL == 0) {
            const int64_t first = b_ * BATCH + (int64_t)tid * R;      // R consecutive particles per thread
#pragma unroll
            for (int k = 0; k < NRAW; k++)
                raw[k] = (first + R <= N) ? __ldcs(reinterpret_cast<const float4 *>(pos + 3 * first) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int64_t i = lo + b_ * BATCH + k * NT + tid;
                raw[k] = i < hi ? __ldcs(tmp + i) : make_float4(0.f, 0.f, 0.f, -1.f);
            }
        }
    };
    if (bi < nbatch) load(bi);
    for (; bi < nbatch; bi += bstep) {
        float4 rec[R];
        int key[R];
        uint32_t pending = 0;
        if (LEVEL == 0) {
            const float *c = reinterpret_cast<const float *>(raw);
            const int64_t first = bi * BATCH + (int64_t)tid * R;
#pragma unroll
            for (int k = 0; k < R; k++) {
                rec[k] = make_float4(c[3 * k], c[3 * k + 1], c[3 * k + 2], 1.0f);
                if (first + R <= N) pending |= 1u << k;      // (prototype: N is a multiple of R)
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; k++) {
                rec[k] = raw[k];
                if (lo + bi * BATCH + k * NT + tid < hi) pending |= 1u << k;
            }
        }
        if (bi + bstep < nbatch) load(bi + bstep);
#pragma unroll
        for (int k = 0; k < R; k++) {
            const uint32_t t = tile_of(g, rec[k].x, rec[k].y, rec[k].z);
            key[k] = LEVEL == 0 ? (int)(t / (uint32_t)D) : (int)t - t0;
        }
        int any;
        do {
            uint32_t p[R];
#pragma unroll
            for (int k = 0; k < R; k++) p[k] = (pending >> k & 1u) ? atomicAdd(&cnt[key[k]], 1u) : 0xffffffffu;
#pragma unroll
            for (int k = 0; k < R; k++)
                if (p[k] < (uint32_t)C) {
                    stage[p[k] * pitch + key[k]] = rec[k];
                    pending &= ~(1u << k);
                }
            any = __syncthreads_or(pending != 0);
            uint32_t *lcur = bE + 1024;      // MODE 3: CTA-private cursors (timing experiment only: positions are not exact)
            for (int b = tid; b < nb; b += NT) {
                const uint32_t c = min(cnt[b], (uint32_t)C);
                const uint32_t n = (c >= (uint32_t)FMIN || c == (uint32_t)C) ? (c & ~1u) : 0u;
                if (n) {
                    uint32_t base;
                    if (MODE >= 4) {      // experiment: every bucket cycles through its own window of 2^(2*MODE+4) records, windows back to back
                        const uint32_t W = 1u << (2 * MODE + 4);
                        base = ((uint32_t)b * W + lcur[b] % (W - 32u)) & ~1u;
                        lcur[b] += n;
                    } else if (MODE == 3) { const uint32_t S_ = bS[b], len = max(bE[b] - S_, 64u) - 32u; base = (S_ + lcur[b] % len) & ~1u; lcur[b] += n; }
                    else base = atomicAdd(&cursors[(LEVEL == 0 ? 0 : t0) + b], n);
                    flush_bucket<MODE>(stage, pitch, b, n, base, bS[b], bE[b], out);
                    if (c & 1u) stage[b] = stage[n * pitch + b];
                }
                cnt[b] = c - n;
            }
            __syncthreads();
        } while (any);
    }
    // drain: whatever is left in the queues
    for (int b = tid; b < nb; b += NT) {
        const uint32_t c = cnt[b];
        if (c) {
            const uint32_t base = atomicAdd(&cursors[(LEVEL == 0 ? 0 : t0) + b], c);
            flush_bucket<MODE>(stage, pitch, b, c, base, bS[b], bE[b], out);
        }
    }
}

// one-pass reference (today's kernel): per-record cursor + 16-byte store
__global__ void __launch_bounds__(256) onepass_kernel(const float *__restrict__ pos, int64_t N, Geo g, uint32_t *__restrict__ cur, float4 *__restrict__ out)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < N; q += (int64_t)gridDim.x * blockDim.x) {
        const float4 *p4 = reinterpret_cast<const float4 *>(pos + 12 * q);
        const float4 a = __ldcs(p4), b = __ldcs(p4 + 1), d = __ldcs(p4 + 2);
        const float c[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
        uint32_t slot[4];
#pragma unroll
        for (int k = 0; k < 4; k++) slot[k] = atomicAdd(&cur[tile_of(g, c[3 * k], c[3 * k + 1], c[3 * k + 2])], 1u);
#pragma unroll
        for (int k = 0; k < 4; k++) out[slot[k]] = make_float4(c[3 * k], c[3 * k + 1], c[3 * k + 2], 1.0f);
    }
}

// occupancy variants of the one-pass scatter: PER particles per thread, MINB resident CTAs of 256 threads asked for
template <int PER, int MINB>
__global__ void __launch_bounds__(256, MINB) onepass_occ_kernel(const float *__restrict__ pos, int64_t N, Geo g, uint32_t *__restrict__ cur, float4 *__restrict__ out)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * PER < N; q += (int64_t)gridDim.x * blockDim.x) {
        float c[3 * PER];
        if (PER == 4) {
            const float4 *p4 = reinterpret_cast<const float4 *>(pos + 12 * q);
            const float4 a = __ldcs(p4), b = __ldcs(p4 + 1), d = __ldcs(p4 + 2);
            const float t[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
#pragma unroll
            for (int k = 0; k < 3 * PER; k++) c[k] = t[k];
        } else {
#pragma unroll
            for (int k = 0; k < 3 * PER; k++) c[k] = __ldcs(pos + 3 * PER * q + k);
        }
        uint32_t slot[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) slot[k] = atomicAdd(&cur[tile_of(g, c[3 * k], c[3 * k + 1], c[3 * k + 2])], 1u);
#pragma unroll
        for (int k = 0; k < PER; k++) out[slot[k]] = make_float4(c[3 * k], c[3 * k + 1], c[3 * k + 2], 1.0f);
    }
}

// one-pass scatter whose record stores carry an L2 evict_last policy (keep the open write frontier resident) while the
// position loads stay evict_first
template <int PRIO>
__global__ void __launch_bounds__(256) onepass_policy_kernel(const float *__restrict__ pos, int64_t N, Geo g, uint32_t *__restrict__ cur, float4 *__restrict__ out)
{
    uint64_t pol;
    if (PRIO == 0) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_unchanged.b64 %0, 0.5;" : "=l"(pol));
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < N; q += (int64_t)gridDim.x * blockDim.x) {
        const float4 *p4 = reinterpret_cast<const float4 *>(pos + 12 * q);
        const float4 a = __ldcs(p4), b = __ldcs(p4 + 1), d = __ldcs(p4 + 2);
        const float c[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
        uint32_t slot[4];
#pragma unroll
        for (int k = 0; k < 4; k++) slot[k] = atomicAdd(&cur[tile_of(g, c[3 * k], c[3 * k + 1], c[3 * k + 2])], 1u);
#pragma unroll
        for (int k = 0; k < 4; k++)
            asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(out + slot[k]), "f"(c[3 * k]), "f"(c[3 * k + 1]), "f"(c[3 * k + 2]),
                         "f"(1.0f), "l"(pol)
                         : "memory");
    }
}

__global__ void verify_kernel(const float4 *__restrict__ out, const uint32_t *__restrict__ starts, Geo g, unsigned long long *__restrict__ res)
{
    unsigned long long bad = 0, sum = 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < g.ntiles; t += gridDim.x * blockDim.x)
        for (uint32_t i = starts[t]; i < starts[t + 1]; i++) {
            const float4 r = out[i];
            bad += tile_of(g, r.x, r.y, r.z) != t || r.w != 1.0f;
            sum += (unsigned long long)__float_as_uint(r.x) + 3ull * __float_as_uint(r.y) + 7ull * __float_as_uint(r.z);
        }
    atomicAdd(&res[0], bad);
    atomicAdd(&res[1], sum);
}

__global__ void possum_kernel(const float *__restrict__ pos, int64_t N, unsigned long long *__restrict__ res)
{
    unsigned long long sum = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        sum += (unsigned long long)__float_as_uint(pos[3 * i]) + 3ull * __float_as_uint(pos[3 * i + 1]) + 7ull * __float_as_uint(pos[3 * i + 2]);
    atomicAdd(&res[2], sum);
}

template <int NT, int R, int FMIN, int MODE = 0>
int run_variant(const char *name, const float *pos, int64_t N, Geo g, const uint32_t *starts, uint32_t *cur_fine, uint32_t *cur_coarse, float4 *tmp,
                float4 *out, unsigned long long *res, int D, int ncoarse, int slots, int num_sms)
{
    const size_t smem = (size_t)slots * 16 + 4 * 1024 * 4;
    auto k0 = split_kernel<0, NT, R, FMIN, MODE>;
    auto k1 = split_kernel<1, NT, R, FMIN, MODE>;
    CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int occ = NT == 1024 ? 1 : 2;
    int parts = (num_sms * occ * 4 + ncoarse - 1) / ncoarse;
    cudaEvent_t e[4];
    for (auto &x : e) cudaEventCreate(&x);
    float tA = 0, tB = 0;
    for (int rep = 0; rep < 2; rep++) {
        cursors_kernel<<<1024, 256>>>(starts, g.ntiles, D, cur_fine, cur_coarse);
        cudaEventRecord(e[0]);
        k0<<<num_sms * occ, NT, smem>>>(pos, nullptr, N, g, starts, cur_coarse, MODE ? out : tmp, D, ncoarse, 1, slots);      // experiments keep tmp intact
        cudaEventRecord(e[1]);
        k1<<<ncoarse * parts, NT, smem>>>(nullptr, tmp, N, g, starts, cur_fine, out, D, ncoarse, parts, slots);
        cudaEventRecord(e[2]);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&tA, e[0], e[1]);
        cudaEventElapsedTime(&tB, e[1], e[2]);
    }
    CK(cudaMemset(res, 0, 16));
    verify_kernel<<<1024, 256>>>(out, starts, g, res);
    unsigned long long h[3];
    CK(cudaMemcpy(h, res, 24, cudaMemcpyDeviceToHost));
    const double sc = 1e9 / (double)N;
    printf("%-34s coarse %6.2f  fine %6.2f  sum %6.2f ms/1e9   (parts %d, slots %d)  bad %llu checksum %s\n", name, tA * sc, tB * sc, (tA + tB) * sc, parts,
           slots, h[0], h[1] == h[2] ? "ok" : "MISMATCH");
    return 0;
}

int main(int argc, char **argv)
{
    const int lg = argc > 2 ? atoi(argv[2]) : 28;
    const int64_t N = (int64_t)1 << lg;
    const int n = 1024;
    Geo g;
    g.n = n; g.nty = n / TY; g.ntz = (n + TZ - 1) / TZ; g.ntiles = (uint32_t)(n / TX) * g.nty * g.ntz; g.fn = (float)n;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int num_sms = prop.multiProcessorCount;
    float *pos; float4 *tmp, *out; uint32_t *counts, *starts, *cur_fine, *cur_coarse; unsigned long long *res;
    CK(cudaMalloc(&pos, N * 12)); CK(cudaMalloc(&tmp, N * 16)); CK(cudaMalloc(&out, N * 16));
    CK(cudaMalloc(&counts, (g.ntiles + 1) * 4)); CK(cudaMalloc(&starts, (g.ntiles + 1) * 4)); CK(cudaMalloc(&cur_fine, (g.ntiles + 1) * 4));
    CK(cudaMalloc(&cur_coarse, 4096)); CK(cudaMalloc(&res, 32));
    fill_pos<<<num_sms * 8, 256>>>(pos, N * 3);
    CK(cudaMemset(counts, 0, (g.ntiles + 1) * 4));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    hist_kernel<<<num_sms * 16, 256>>>(pos, N, g, counts);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float th; cudaEventElapsedTime(&th, a, b);
    void *d_tmp = nullptr; size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(d_tmp, tb, counts, starts, g.ntiles + 1);
    CK(cudaMalloc(&d_tmp, tb));
    cub::DeviceScan::ExclusiveSum(d_tmp, tb, counts, starts, g.ntiles + 1);
    CK(cudaMemset(res, 0, 32));
    possum_kernel<<<num_sms * 8, 256>>>(pos, N, res);
    CK(cudaDeviceSynchronize());
    const double sc = 1e9 / (double)N;
    printf("N = 2^%d, mesh %d^3, %u tiles; histogram (global atomics) %.2f ms/1e9\n", lg, n, g.ntiles, th * sc);

    // today's one-pass scatter
    {
        float t1 = 0;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemcpy(cur_fine, starts, (g.ntiles + 1) * 4, cudaMemcpyDeviceToDevice));
            cudaEventRecord(a);
            onepass_kernel<<<num_sms * 16, 256>>>(pos, N, g, cur_fine, out);
            cudaEventRecord(b);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&t1, a, b);
        }
        CK(cudaMemset(res, 0, 16));
        verify_kernel<<<1024, 256>>>(out, starts, g, res);
        unsigned long long h[3];
        CK(cudaMemcpy(h, res, 24, cudaMemcpyDeviceToHost));
        printf("%-34s %6.2f ms/1e9  bad %llu checksum %s\n", "one pass, 16-B stores", t1 * sc, h[0], h[1] == h[2] ? "ok" : "MISMATCH");
    }
    {
        auto run1 = [&](const char *name, auto kern, int per_sm) -> int {
            float t1 = 0;
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaMemcpy(cur_fine, starts, (g.ntiles + 1) * 4, cudaMemcpyDeviceToDevice));
                cudaEventRecord(a);
                kern<<<num_sms * per_sm, 256>>>(pos, N, g, cur_fine, out);
                cudaEventRecord(b);
                CK(cudaDeviceSynchronize());
                cudaEventElapsedTime(&t1, a, b);
            }
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, 0);
            printf("%-34s %6.2f ms/1e9  (%d CTAs of 256 resident per SM)\n", name, t1 * sc, nb);
            return 0;
        };
        if (run1("one pass, 4/thread, minb 4", onepass_occ_kernel<4, 4>, 16)) return 1;
        if (run1("one pass, 4/thread, minb 6", onepass_occ_kernel<4, 6>, 24)) return 1;
        if (run1("one pass, 4/thread, minb 8", onepass_occ_kernel<4, 8>, 32)) return 1;
        if (run1("one pass, 2/thread, minb 8", onepass_occ_kernel<2, 8>, 32)) return 1;
        if (run1("one pass, 1/thread, minb 8", onepass_occ_kernel<1, 8>, 32)) return 1;
        if (run1("one pass, 2/thread, minb 4", onepass_occ_kernel<2, 4>, 16)) return 1;
        if (run1("one pass, stores L2 evict_last", onepass_policy_kernel<0>, 16)) return 1;
        if (run1("one pass, stores evict_last 50 %", onepass_policy_kernel<1>, 16)) return 1;
        if (argc > 1 && argv[1][0] == 'o') return 0;
    }
    const int D = (int)ceil(sqrt((double)g.ntiles));
    const int ncoarse = (g.ntiles + D - 1) / D;
    printf("two-level: D = %d tiles per coarse bucket, %d coarse buckets\n", D, ncoarse);
    CK(cudaMemset(out, 0, N * 16));
    if (run_variant<1024, 4, 2>("1024 thr, R=4, flush>=2", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 12288, num_sms)) return 1;
    if (argc > 1 && argv[1][0] == 'p') return 0;      // profiling run: first variant only
    printf("experiments (results not exact by construction):\n");
    if (run_variant<1024, 4, 2, 1>("  no global stores", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 12288, num_sms)) return 1;
    if (run_variant<1024, 4, 2, 2>("  stores into a 1 MB window", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 12288, num_sms)) return 1;
    if (run_variant<1024, 4, 2, 3>("  CTA-private cursors", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 12288, num_sms)) return 1;
    if (run_variant<1024, 4, 2, 4>("  private, 64 KB window per bucket", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 12288, num_sms)) return 1;
    if (run_variant<1024, 4, 2, 5>("  private, 256 KB window per bucket", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 12288, num_sms)) return 1;
    if (run_variant<1024, 4, 2, 6>("  private, 1 MB window per bucket", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 12288, num_sms)) return 1;
    if (run_variant<1024, 4, 2, 7>("  private, 4 MB window per bucket", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 12288, num_sms)) return 1;
    if (run_variant<512, 4, 2, 3>("  CTA-private cursors, 512 thr x2", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 6144, num_sms)) return 1;
    if (run_variant<1024, 4, 8, 3>("  CTA-private cursors, flush>=8", pos, N, g, starts, cur_fine, cur_coarse, tmp, out, res, D, ncoarse, 13056, num_sms)) return 1;
    return 0;
}
