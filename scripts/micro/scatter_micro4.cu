// Micro-benchmark 4 (round 2): is the bucket scatter bound by address-translation / page locality?
// Frontier pattern (per-tile cursor + 16-byte record store at tile*cap + slot) where, at any time, all CTAs scatter into a
// WINDOW of consecutive tiles (what the second level of a coarse-then-fine bucketing does), for several window sizes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/scatter_micro4 scripts/micro/scatter_micro4.cu && timeout 120 /tmp/scatter_micro4
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

// particles are processed in index order (grid-stride); particle group g belongs to window g / groups_per_window
template <bool LOAD>
__global__ void __launch_bounds__(256) frontier_kernel(const float4 *__restrict__ src, float4 *__restrict__ rec, int64_t N, uint32_t ntiles, uint32_t cap,
                                                       uint32_t *__restrict__ cur, uint32_t wtiles)
{
    const int64_t ngroups = N / 4;
    const int64_t nwin = ntiles / wtiles;
    const int64_t gpw = (ngroups + nwin - 1) / nwin;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w0 = (uint32_t)(g / gpw) * wtiles;
        uint32_t tile[4], slot[4];
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            tile[q] = w0 + hash((uint32_t)(4 * g + q)) % wtiles;
            v[q] = LOAD ? __ldcs(src + 4 * g + q) : make_float4((float)g, 2.f, 3.f, 1.f);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) slot[q] = atomicAdd(&cur[tile[q]], 1u);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (slot[q] >= cap) continue;
            rec[(size_t)tile[q] * cap + slot[q]] = v[q];
        }
    }
}

int main()
{
    const int64_t N = (int64_t)1 << 28;
    const uint32_t nt = 1u << 19;
    const uint32_t cap = (uint32_t)(N / nt) + (uint32_t)(8 * sqrt((double)N / nt)) + 16;
    float4 *rec, *src;
    cudaMalloc(&rec, (size_t)nt * cap * 16);
    cudaMalloc(&src, (size_t)N * 16); cudaMemset(src, 0, (size_t)N * 16);
    uint32_t *cur; cudaMalloc(&cur, nt * 4);
    const int blocks = 148 * 16;
    const double sc = 1e9 / (double)N;
    for (uint32_t wt : {nt, nt / 8, nt / 64, nt / 512, nt / 4096}) {
        float a = timeit([&] { cudaMemsetAsync(cur, 0, nt * 4); frontier_kernel<false><<<blocks, 256>>>(src, rec, N, nt, cap, cur, wt); });
        float b = timeit([&] { cudaMemsetAsync(cur, 0, nt * 4); frontier_kernel<true><<<blocks, 256>>>(src, rec, N, nt, cap, cur, wt); });
        printf("J window %6u tiles (%7.1f MB of records): atomic + st.v4 %.2f ms/1e9   with 16-B record loads %.2f\n", wt, (double)wt * cap * 16 / 1048576.0, a * sc, b * sc);
    }
    return 0;
}
