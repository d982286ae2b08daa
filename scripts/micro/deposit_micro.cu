// Micro-benchmarks behind the round-2 deposit / bucketing design (B200, sm_100a).  Every number is a CUDA-event time of
// the second launch.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/deposit_micro scripts/micro/deposit_micro.cu && timeout 120 /tmp/deposit_micro
//
//   A  list building in shared memory, per particle: ATOMS.EXCH / ATOMS.ADD (with return) vs a plain 16-bit store
//      followed by a read-back ("store race": the lane whose id survives owns the slot)
//   B  emission of finished tile rows straight into the grid: every warp walks the x-planes of its (y-rows, 32 z) column
//      of an 8x8x32 tile and issues R coalesced 32-float reductions (+ R two-lane halo reductions) per plane, addresses
//      laid out like the real mesh (1024 x 1024 x 1026 floats).  R = 3 rows per 1 cell row, 4 per 2, 6 per 4.
//   C  bucket scatter: one returning global atomic on a per-tile cursor + one scattered record store per particle,
//      as a function of the number of tiles and the record size
//   D  the FMA body of the per-cell loop at different lane efficiencies (upper bound of the accumulate phase)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <class F>
float timeit(F launch)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return ms;
}

// ---------------------------------------------------------------------------------------------------- A
// MODE 0: atomicExch u32   1: atomicAdd u32 (returning)   2: u16 store + barrier + read-back   3: u16 store only
template <int MODE>
__global__ void __launch_bounds__(256) list_kernel(uint32_t *sink, int iters, int ncell)
{
    extern __shared__ uint32_t s[];
    uint16_t *s16 = reinterpret_cast<uint16_t *>(s);
    for (int i = threadIdx.x; i < ncell; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint32_t acc = 0, key = hash(blockIdx.x * 256 + threadIdx.x);
    for (int it = 0; it < iters; it++) {
        uint32_t c[8];
#pragma unroll
        for (int q = 0; q < 8; q++) { key = hash(key + q); c[q] = key % ncell; }
        if (MODE == 0) {
#pragma unroll
            for (int q = 0; q < 8; q++) acc += atomicExch(&s[c[q]], key + q);
        } else if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < 8; q++) acc += atomicAdd(&s[c[q]], 1u);
        } else {
            const uint16_t me = (uint16_t)(threadIdx.x * 8 + it);
#pragma unroll
            for (int q = 0; q < 8; q++) s16[c[q]] = (uint16_t)(me + q);
            if (MODE == 2) {
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 8; q++) acc += (s16[c[q]] == (uint16_t)(me + q));
                __syncthreads();
            }
        }
    }
    if (acc == 0xdeadbeefu) sink[0] = acc;
}

// ---------------------------------------------------------------------------------------------------- B
// grid n x n x ldz; tile 8 x 8 x 32; CTA = tile, warp w owns rows [w*Y, w*Y+Y) (8/Y warps); per x-plane (10 of them)
// the warp adds Y+2 rows of 32 floats (+ the two z-halo cells, one 2-lane RED per row if HALO).
template <int Y, bool HALO, bool V2>
__global__ void __launch_bounds__(256) emit_kernel(float *__restrict__ grid, int n, int64_t ldz)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w >= 8 / Y) return;
    const int ntz = n / 32, nty = n / 8;
    const int tile = blockIdx.x;
    const int tz = tile % ntz, ty = (tile / ntz) % nty, tx = tile / (ntz * nty);
    const int x0 = tx * 8, y0 = ty * 8 + w * Y, z0 = tz * 32;
    const int64_t sx = (int64_t)n * ldz;
    float v = 1.0f + lane;
    for (int px = -1; px <= 8; px++) {
        const int gx = (x0 + px + n) % n;
#pragma unroll
        for (int r = 0; r < Y + 2; r++) {
            const int gy = (y0 + r - 1 + n) % n;
            float *row = grid + gx * sx + (int64_t)gy * ldz;
            if (V2) {
                if (lane < 16) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(row + z0 + 2 * lane), "f"(v), "f"(v) : "memory");
            } else {
                atomicAdd(row + z0 + lane, v);
            }
            if (HALO && lane < 2) atomicAdd(row + (lane ? (z0 + 32) % n : (z0 - 1 + n) % n), v);
            v += 1.0f;
        }
    }
}

// same traffic through a shared-memory output tile (what round 1 does): tile+halo 10x10x34 flushed by 8 warps
__global__ void __launch_bounds__(256) flush_kernel(float *__restrict__ grid, int n, int64_t ldz)
{
    __shared__ float out[10 * 10 * 34];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 3400; i += 256) out[i] = 1.0f;
    __syncthreads();
    const int ntz = n / 32, nty = n / 8;
    const int tile = blockIdx.x;
    const int tz = tile % ntz, ty = (tile / ntz) % nty, tx = tile / (ntz * nty);
    const int x0 = tx * 8, y0 = ty * 8, z0 = tz * 32;
    const int64_t sx = (int64_t)n * ldz;
    for (int oy = w; oy < 10; oy += 8) {
        const int gy = (y0 + oy - 1 + n) % n;
        for (int ox = 0; ox < 10; ox++) {
            const int gx = (x0 + ox - 1 + n) % n;
            float *row = grid + gx * sx + (int64_t)gy * ldz;
            const float *r = out + (ox * 10 + oy) * 34;
            atomicAdd(row + z0 + lane, r[lane + 1]);
            if (lane < 2) atomicAdd(row + (lane ? (z0 + 32) % n : (z0 - 1 + n) % n), r[lane ? 33 : 0]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------- C
// particles i = 0..N-1 with pseudo-random tile; cursor[tile] hands out slots in [tile*cap, (tile+1)*cap)
template <int RECB, bool LOADPOS>
__global__ void __launch_bounds__(256) scatter_kernel(const float *__restrict__ pos, int64_t N, uint32_t ntiles, uint32_t cap,
                                                      uint32_t *__restrict__ cursor, unsigned char *__restrict__ rec)
{
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g * 4 < N; g += (int64_t)gridDim.x * blockDim.x) {
        float c[12];
        if (LOADPOS) {
            const float4 *p4 = reinterpret_cast<const float4 *>(pos + 12 * g);
            const float4 a = __ldcs(p4), b = __ldcs(p4 + 1), d = __ldcs(p4 + 2);
            c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
            c[8] = d.x; c[9] = d.y; c[10] = d.z; c[11] = d.w;
        } else {
#pragma unroll
            for (int q = 0; q < 12; q++) c[q] = (float)q;
        }
        uint32_t tile[4], slot[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t h = hash((uint32_t)(4 * g + q));
            if (LOADPOS) h ^= __float_as_uint(c[3 * q]) & 1u;
            tile[q] = h % ntiles;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) slot[q] = atomicAdd(&cursor[tile[q]], 1u);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (slot[q] >= cap) continue;
            const size_t o = ((size_t)tile[q] * cap + slot[q]) * RECB;
            if (RECB == 16) *reinterpret_cast<float4 *>(rec + o) = make_float4(c[3 * q], c[3 * q + 1], c[3 * q + 2], 1.0f);
            else *reinterpret_cast<float2 *>(rec + o) = make_float2(c[3 * q], c[3 * q + 1]);
        }
    }
}

// histogram only (no-return reductions), the first pass of the exact bucketing
__global__ void __launch_bounds__(256) hist_kernel(int64_t N, uint32_t ntiles, uint32_t *__restrict__ counts)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&counts[hash((uint32_t)i) % ntiles], 1u);
}

// ---------------------------------------------------------------------------------------------------- D
// per-cell accumulate loop: every lane runs `cnt` iterations (cnt from a Poisson-like table), 27 FMAs + weights each,
// records read from shared memory at random positions
__global__ void __launch_bounds__(256) accum_kernel(float *sink, int steps, float mean)
{
    __shared__ float4 srec[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) srec[i] = make_float4(0.1f * (i & 7) - 0.4f, 0.05f * (i & 15) - 0.4f, 0.02f * (i & 31) - 0.3f, 1.0f);
    __syncthreads();
    float S[3][3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int c = 0; c < 3; c++) S[a][b][c] = 0.0f;
    uint32_t key = hash(blockIdx.x * 256 + threadIdx.x);
    for (int st = 0; st < steps; st++) {
        key = hash(key);
        // Poisson(mean) by inversion on a uniform
        float u = (key >> 8) * (1.0f / 16777216.0f), p = __expf(-mean), cdf = p;
        int cnt = 0;
        while (u > cdf && cnt < 12) { cnt++; p *= mean / cnt; cdf += p; }
        uint32_t i = key & 2047;
        for (int k = 0; k < cnt; k++) {
            const float4 r = srec[i];
            i = (i * 5 + 1) & 2047;
            float wx[3], wy[3], wz[3];
            { const float a = 0.5f + r.x, b = 0.5f - r.x; wx[0] = 0.5f * a * a; wx[1] = 0.75f - r.x * r.x; wx[2] = 0.5f * b * b; }
            { const float a = 0.5f + r.y, b = 0.5f - r.y; wy[0] = 0.5f * a * a * r.w; wy[1] = (0.75f - r.y * r.y) * r.w; wy[2] = 0.5f * b * b * r.w; }
            { const float a = 0.5f + r.z, b = 0.5f - r.z; wz[0] = 0.5f * a * a; wz[1] = 0.75f - r.z * r.z; wz[2] = 0.5f * b * b; }
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float t = wy[b] * wz[c];
                    S[0][b][c] = fmaf(wx[0], t, S[0][b][c]);
                    S[1][b][c] = fmaf(wx[1], t, S[1][b][c]);
                    S[2][b][c] = fmaf(wx[2], t, S[2][b][c]);
                }
        }
    }
    float tot = 0;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int c = 0; c < 3; c++) tot += S[a][b][c];
    if (tot == 1234.5f) sink[0] = tot;
}

int main()
{
    uint32_t *sink; cudaMalloc(&sink, 64);
    // ---- A
    {
        const int blocks = 148 * 4, iters = 256;
        const double ops = (double)blocks * 256 * iters * 8;
        for (int nc : {2048, 4096}) {
            float t0 = timeit([&] { list_kernel<0><<<blocks, 256, nc * 4>>>(sink, iters, nc); });
            float t1 = timeit([&] { list_kernel<1><<<blocks, 256, nc * 4>>>(sink, iters, nc); });
            float t2 = timeit([&] { list_kernel<2><<<blocks, 256, nc * 4>>>(sink, iters, nc); });
            float t3 = timeit([&] { list_kernel<3><<<blocks, 256, nc * 4>>>(sink, iters, nc); });
            printf("A ncell %4d: atomicExch %.1f Gop/s  atomicAdd(ret) %.1f  store+barrier+readback %.1f  store only %.1f   (ms/1e9: %.2f %.2f %.2f %.2f)\n",
                   nc, ops / t0 / 1e6, ops / t1 / 1e6, ops / t2 / 1e6, ops / t3 / 1e6, t0 * 1e9 / ops, t1 * 1e9 / ops, t2 * 1e9 / ops, t3 * 1e9 / ops);
        }
    }
    // ---- B
    {
        const int n = 1024;
        const int64_t ldz = 1026;
        float *grid; cudaMalloc(&grid, (size_t)n * n * ldz * 4); cudaMemset(grid, 0, (size_t)n * n * ldz * 4);
        const int ntiles = (n / 8) * (n / 8) * (n / 32);
        float a1 = timeit([&] { emit_kernel<1, true, false><<<ntiles, 256>>>(grid, n, ldz); });
        float a1n = timeit([&] { emit_kernel<1, false, false><<<ntiles, 256>>>(grid, n, ldz); });
        float a2 = timeit([&] { emit_kernel<2, true, false><<<ntiles, 128>>>(grid, n, ldz); });
        float a4 = timeit([&] { emit_kernel<4, true, false><<<ntiles, 64>>>(grid, n, ldz); });
        float v1 = timeit([&] { emit_kernel<1, true, true><<<ntiles, 256>>>(grid, n, ldz); });
        float fl = timeit([&] { flush_kernel<<<ntiles, 256>>>(grid, n, ldz); });
        printf("B emission of a 1024^3 mesh (8x8x32 tiles, 10 planes): Y=1 3 rows/plane %.2f ms (no z-halo REDs %.2f)   Y=2 %.2f   Y=4 %.2f   Y=1 RED.v2 %.2f   smem-tile flush %.2f\n",
               a1, a1n, a2, a4, v1, fl);
        cudaFree(grid);
    }
    // ---- C
    {
        const int64_t N = (int64_t)1 << 28;
        float *pos; cudaMalloc(&pos, N * 12); cudaMemset(pos, 0, N * 12);
        for (uint32_t nt : {1u << 15, 1u << 17, 1u << 19}) {
            const uint32_t cap = (uint32_t)(N / nt) + (uint32_t)(8 * sqrt((double)N / nt)) + 16;
            uint32_t *cur; cudaMalloc(&cur, nt * 4);
            unsigned char *rec; cudaMalloc(&rec, (size_t)nt * cap * 16);
            const int blocks = 148 * 16;
            auto run = [&](auto k) { return timeit([&] { cudaMemsetAsync(cur, 0, nt * 4); k(); }); };
            float h = run([&] { hist_kernel<<<blocks, 256>>>(N, nt, cur); });
            float s16 = run([&] { scatter_kernel<16, false><<<blocks, 256>>>(pos, N, nt, cap, cur, rec); });
            float s16p = run([&] { scatter_kernel<16, true><<<blocks, 256>>>(pos, N, nt, cap, cur, rec); });
            float s8p = run([&] { scatter_kernel<8, true><<<blocks, 256>>>(pos, N, nt, cap, cur, rec); });
            const double sc = 1e9 / (double)N;
            printf("C tiles %7u (cap %u): hist %.2f ms/1e9   scatter 16B (no loads) %.2f   16B + pos loads %.2f   8B + pos loads %.2f\n",
                   nt, cap, h * sc, s16 * sc, s16p * sc, s8p * sc);
            cudaFree(cur); cudaFree(rec);
        }
        cudaFree(pos);
    }
    // ---- D
    {
        const int blocks = 148 * 6, steps = 4096;
        for (float mean : {0.93f, 1.86f}) {
            float t = timeit([&] { accum_kernel<<<blocks, 256>>>((float *)sink, steps, mean); });
            const double parts = (double)blocks * 256 * steps * mean;
            printf("D accumulate loop, Poisson mean %.2f per lane-step: %.1f Gpart/s  (%.2f ms per 1e9 particles)\n", mean, parts / t / 1e6, t * 1e9 / parts);
        }
    }
    return 0;
}
