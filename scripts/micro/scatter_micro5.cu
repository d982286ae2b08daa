// Micro-benchmark 5 (round 2): are the bucket scatter's partial-sector stores paying a DRAM fill per sector?
// Frontier pattern (per-tile cursor + record store at tile*cap + slot, 2^19 tiles, 2^28 particles) with
//   16-byte records, st.global.v4.f32                        (today: two records share a 32-byte sector)
//   32-byte records, two st.global.v4.f32 to the same sector
//   32-byte records, ONE st.global.v8.f32 (256-bit store, sm_100): a full sector per store, no fill needed
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/scatter_micro5 scripts/micro/scatter_micro5.cu && timeout 120 /tmp/scatter_micro5
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

template <int MODE>
__global__ void __launch_bounds__(256) frontier_kernel(const float *__restrict__ pos, unsigned char *__restrict__ rec, int64_t N, uint32_t ntiles, uint32_t cap,
                                                       uint32_t *__restrict__ cur)
{
    constexpr int RB = MODE == 0 ? 16 : 32;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g * 4 < N; g += (int64_t)gridDim.x * blockDim.x) {
        const float4 *p4 = reinterpret_cast<const float4 *>(pos + 12 * g);
        const float4 a = __ldcs(p4), b = __ldcs(p4 + 1), d = __ldcs(p4 + 2);
        const float c[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
        uint32_t tile[4], slot[4];
#pragma unroll
        for (int q = 0; q < 4; q++) tile[q] = (hash((uint32_t)(4 * g + q)) ^ (__float_as_uint(c[3 * q]) & 1u)) % ntiles;
#pragma unroll
        for (int q = 0; q < 4; q++) slot[q] = atomicAdd(&cur[tile[q]], 1u);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (slot[q] >= cap) continue;
            unsigned char *p = rec + ((size_t)tile[q] * cap + slot[q]) * RB;
            const float x = c[3 * q], y = c[3 * q + 1], z = c[3 * q + 2];
            if (MODE == 0) {
                *reinterpret_cast<float4 *>(p) = make_float4(x, y, z, 1.0f);
            } else if (MODE == 1) {
                *reinterpret_cast<float4 *>(p) = make_float4(x, y, z, 1.0f);
                *reinterpret_cast<float4 *>(p + 16) = make_float4(z, y, x, 2.0f);
            } else {
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(1.0f), "f"(z), "f"(y),
                             "f"(x), "f"(2.0f)
                             : "memory");
            }
        }
    }
}

int main()
{
    const int64_t N = (int64_t)1 << 28;
    const uint32_t nt = 1u << 19;
    const uint32_t cap = (uint32_t)(N / nt) + (uint32_t)(8 * sqrt((double)N / nt)) + 16;
    float *pos; cudaMalloc(&pos, N * 12); cudaMemset(pos, 0, N * 12);
    unsigned char *rec; cudaMalloc(&rec, (size_t)nt * cap * 32);
    uint32_t *cur; cudaMalloc(&cur, nt * 4);
    const int blocks = 148 * 16;
    const double sc = 1e9 / (double)N;
    float t0 = timeit([&] { cudaMemsetAsync(cur, 0, nt * 4); frontier_kernel<0><<<blocks, 256>>>(pos, rec, N, nt, cap, cur); });
    float t1 = timeit([&] { cudaMemsetAsync(cur, 0, nt * 4); frontier_kernel<1><<<blocks, 256>>>(pos, rec, N, nt, cap, cur); });
    float t2 = timeit([&] { cudaMemsetAsync(cur, 0, nt * 4); frontier_kernel<2><<<blocks, 256>>>(pos, rec, N, nt, cap, cur); });
    printf("K 2^19 tiles, pos loads + cursor atomic + record store: 16-B st.v4 %.2f ms/1e9   32-B as 2 x st.v4 %.2f   32-B st.v8 %.2f\n", t0 * sc, t1 * sc, t2 * sc);
    return 0;
}
