// Micro-benchmark 3 (round 2): cost of ONE scattered 16-byte record write per particle, by instruction.
// Slots are unique and pseudo-random over a 4 GiB record array (no two lanes share a sector at the same time):
//   st.global.v4 (default / .cg / .cs / .wt), 4 x st.global.b32, st.global.v2 (8 B), red.global.add.v4.f32 into zeroed memory,
//   atom.global.exch.b128, and a frontier pattern (tile cursors, slots sequential per tile) for the two best.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/scatter_micro3 scripts/micro/scatter_micro3.cu && timeout 120 /tmp/scatter_micro3
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

// bijective scramble of a 28-bit index (odd multiplier + xorshift): unique slots
__device__ __forceinline__ uint32_t scramble28(uint32_t i)
{
    i = (i * 0x9E3779B1u) & 0x0fffffffu;
    i ^= i >> 13;
    i = (i * 0x85EBCA6Bu) & 0x0fffffffu;
    i ^= i >> 11;
    return i & 0x0fffffffu;
}

template <int MODE>
__global__ void __launch_bounds__(256) wr_kernel(float4 *__restrict__ rec, int64_t N)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = scramble28((uint32_t)i);
        float4 v = make_float4((float)i, 2.f, 3.f, 1.f);
        float4 *p = rec + s;
        if (MODE == 0) *p = v;
        else if (MODE == 1) __stcg(p, v);
        else if (MODE == 2) __stcs(p, v);
        else if (MODE == 3) __stwt(p, v);
        else if (MODE == 4) { float *q = (float *)p; q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w; }
        else if (MODE == 5) { *(float2 *)p = make_float2(v.x, v.y); }
        else if (MODE == 6) { *(float *)p = v.x; }
        else if (MODE == 7) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        else if (MODE == 8) {
            asm volatile("{ .reg .b128 t, o; mov.b128 t, {%1, %2}; atom.global.exch.b128 o, [%0], t; }" ::"l"(p), "l"(((unsigned long long)__float_as_uint(v.y) << 32) | __float_as_uint(v.x)), "l"(((unsigned long long)__float_as_uint(v.w) << 32) | __float_as_uint(v.z)) : "memory");
        }
        else if (MODE == 9) atomicAdd((float *)p, v.x);
    }
}

// frontier pattern: per-tile cursor (returning atomic), record written at tile*cap + slot
template <int MODE>
__global__ void __launch_bounds__(256) frontier_kernel(float4 *__restrict__ rec, int64_t N, uint32_t ntiles, uint32_t cap, uint32_t *__restrict__ cur)
{
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g * 4 < N; g += (int64_t)gridDim.x * blockDim.x) {
        uint32_t tile[4], slot[4];
#pragma unroll
        for (int q = 0; q < 4; q++) tile[q] = hash((uint32_t)(4 * g + q)) % ntiles;
#pragma unroll
        for (int q = 0; q < 4; q++) slot[q] = atomicAdd(&cur[tile[q]], 1u);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (slot[q] >= cap) continue;
            float4 *p = rec + (size_t)tile[q] * cap + slot[q];
            const float4 v = make_float4((float)g, 2.f, 3.f, 1.f);
            if (MODE == 0) *p = v;
            else asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
    }
}

int main()
{
    const int64_t N = (int64_t)1 << 28;
    float4 *rec; const size_t NREC = (size_t)(N * 1.45); cudaMalloc(&rec, NREC * 16); cudaMemset(rec, 0, NREC * 16);
    const int blocks = 148 * 16;
    const double sc = 1e9 / (double)N;
    const char *names[] = {"st.v4", "st.cg.v4", "st.cs.v4", "st.wt.v4", "4 x st.b32", "st.v2 (8 B)", "st.b32 (4 B)", "red.add.v4.f32", "atom.exch.b128", "red.add.f32 (4 B)"};
    float t;
    t = timeit([&] { wr_kernel<0><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[0], t * sc);
    t = timeit([&] { wr_kernel<1><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[1], t * sc);
    t = timeit([&] { wr_kernel<2><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[2], t * sc);
    t = timeit([&] { wr_kernel<3><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[3], t * sc);
    t = timeit([&] { wr_kernel<4><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[4], t * sc);
    t = timeit([&] { wr_kernel<5><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[5], t * sc);
    t = timeit([&] { wr_kernel<6><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[6], t * sc);
    t = timeit([&] { wr_kernel<7><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[7], t * sc);
    t = timeit([&] { wr_kernel<8><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[8], t * sc);
    t = timeit([&] { wr_kernel<9><<<blocks, 256>>>(rec, N); }); printf("H %-18s %.2f ms/1e9\n", names[9], t * sc);
    for (uint32_t nt : {1u << 17, 1u << 19}) {
        const uint32_t cap = (uint32_t)(N / nt) + (uint32_t)(8 * sqrt((double)N / nt)) + 16;
        if ((size_t)nt * cap > NREC) { printf("skip\n"); continue; }
        uint32_t *cur; cudaMalloc(&cur, nt * 4);
        float a = timeit([&] { cudaMemsetAsync(cur, 0, nt * 4); frontier_kernel<0><<<blocks, 256>>>(rec, N, nt, cap, cur); });
        float b = timeit([&] { cudaMemsetAsync(cur, 0, nt * 4); frontier_kernel<1><<<blocks, 256>>>(rec, N, nt, cap, cur); });
        printf("I frontier %u tiles: atomic + st.v4 %.2f ms/1e9   atomic + red.add.v4.f32 %.2f\n", nt, a * sc, b * sc);
        cudaFree(cur);
    }
    return 0;
}
