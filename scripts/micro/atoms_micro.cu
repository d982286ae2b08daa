// Micro-benchmark: the per-lane rates the deposit and the two-level bucketing are built on (B200, sm_100a).
//   A  shared-memory atomicAdd(u32) with return, random addresses among NB counters (multisplit ranking)
//   B  shared-memory atomicExch(u32), random addresses (per-cell list building)
//   C  __match_any_sync on random keys (the alternative way to rank)
//   D  global float reductions to consecutive cells: scalar RED vs red.global.add.v2.f32 (tile flush)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_micro atoms_micro.cu ; run under `timeout 60`.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(256) smem_kernel(uint32_t *sink, int iters, int nb)
{
    extern __shared__ uint32_t s[];
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint32_t acc = 0, key = hash(blockIdx.x * 256 + threadIdx.x);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 16; q++) {
            key = hash(key + q);
            if (MODE == 0) acc += atomicAdd(&s[key % nb], 1u);
            else if (MODE == 1) acc += atomicExch(&s[key % nb], key);
            else acc += __match_any_sync(0xffffffffu, key % nb);
        }
    }
    if (acc == 0xdeadbeefu) sink[0] = acc;
}

template <int V2>
__global__ void __launch_bounds__(256) red_kernel(float *grid, int64_t ncell, int iters)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int it = 0; it < iters; it++) {
        // every warp flushes one 32-cell row; rows are spread over the grid like the tiles' rows are
        const int64_t row = (int64_t)(hash((uint32_t)(warp + it * nwarps)) % (uint32_t)(ncell / 32)) * 32;
        if (V2) {
            if (lane < 16) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(grid + row + 2 * lane), "f"(1.0f), "f"(2.0f) : "memory");
        } else {
            atomicAdd(grid + row + lane, 1.0f);
        }
    }
}

template <class F>
float timeit(F launch)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    uint32_t *sink; cudaMalloc(&sink, 4);
    const int blocks = 148 * 4, iters = 256;
    const double ops = (double)blocks * 256 * iters * 16;
    for (int nb : {128, 512, 1024, 2048}) {
        float t0 = timeit([&] { smem_kernel<0><<<blocks, 256, nb * 4>>>(sink, iters, nb); });
        float t1 = timeit([&] { smem_kernel<1><<<blocks, 256, nb * 4>>>(sink, iters, nb); });
        float t2 = timeit([&] { smem_kernel<2><<<blocks, 256, nb * 4>>>(sink, iters, nb); });
        printf("nb %4d: smem atomicAdd %.1f Gop/s   atomicExch %.1f Gop/s   match_any %.1f Gop/s\n", nb, ops / t0 / 1e6, ops / t1 / 1e6, ops / t2 / 1e6);
    }
    const int64_t ncell = (int64_t)1 << 30;   // 4 GiB grid, like nmesh 1024
    float *grid; cudaMalloc(&grid, ncell * 4); cudaMemset(grid, 0, ncell * 4);
    const int rblocks = 148 * 8, riters = 512;
    const double cells = (double)rblocks * 8 * riters * 32;
    float s = timeit([&] { red_kernel<0><<<rblocks, 256>>>(grid, ncell, riters); });
    float v = timeit([&] { red_kernel<1><<<rblocks, 256>>>(grid, ncell, riters); });
    printf("flush of 32-cell rows: scalar RED %.1f Gcell/s   RED.v2 %.1f Gcell/s\n", cells / s / 1e6, cells / v / 1e6);
    return 0;
}
