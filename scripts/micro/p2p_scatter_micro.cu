// Micro-benchmark 7 (round 2, two GPUs): can the bucket scatter write its 16-byte records STRAIGHT into the owner's memory?
// One process, two devices with peer access.  The frontier pattern of the bucket scatter (per-tile cursor in LOCAL memory,
// record store at cursor position) with a fraction of the tiles living in the peer's memory:
//   0/8, 4/8 (two ranks), 7/8 (eight ranks: everything but the own slab leaves the GPU), 8/8
// and, for reference, a plain coalesced copy kernel to the peer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/p2p_scatter scripts/micro/p2p_scatter_micro.cu && /tmp/p2p_scatter
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// tiles [0, nremote) live in `rem`, the others in `loc`; both buffers hold ntiles * cap records
__global__ void __launch_bounds__(256) frontier_kernel(const float *__restrict__ pos, int64_t N, uint32_t ntiles, uint32_t nremote, uint32_t cap,
                                                       uint32_t *__restrict__ cur, float4 *__restrict__ loc, float4 *__restrict__ rem)
{
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g * 4 < N; g += (int64_t)gridDim.x * blockDim.x) {
        const float4 *p4 = reinterpret_cast<const float4 *>(pos + 12 * g);
        const float4 a = __ldcs(p4), b = __ldcs(p4 + 1), d = __ldcs(p4 + 2);
        const float c[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
        uint32_t tile[4], slot[4];
#pragma unroll
        for (int q = 0; q < 4; q++) tile[q] = (hash32((uint32_t)(4 * g + q)) ^ (__float_as_uint(c[3 * q]) & 1u)) % ntiles;
#pragma unroll
        for (int q = 0; q < 4; q++) slot[q] = atomicAdd(&cur[tile[q]], 1u);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (slot[q] >= cap) continue;
            float4 *dst = (tile[q] < nremote ? rem : loc) + (size_t)tile[q] * cap + slot[q];
            *dst = make_float4(c[3 * q], c[3 * q + 1], c[3 * q + 2], 1.0f);
        }
    }
}

__global__ void __launch_bounds__(256) copy_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = __ldcs(src + i);
}

int main()
{
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs two GPUs\n"); return 0; }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    printf("peer access 0 -> 1: %d\n", can);
    const int64_t N = (int64_t)1 << 27;
    const uint32_t nt = 1u << 19;
    const uint32_t cap = (uint32_t)(N / nt) + 8 * 16 + 16;
    float4 *rem;
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&rem, (size_t)nt * cap * 16));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    float *pos; float4 *loc; uint32_t *cur;
    CK(cudaMalloc(&pos, N * 12)); CK(cudaMemset(pos, 0, N * 12));
    CK(cudaMalloc(&loc, (size_t)nt * cap * 16));
    CK(cudaMalloc(&cur, nt * 4));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const double sc = 1e9 / (double)N;
    for (int eighths : {0, 4, 7, 8}) {
        float ms = 0;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemsetAsync(cur, 0, nt * 4));
            cudaEventRecord(a);
            frontier_kernel<<<148 * 16, 256>>>(pos, N, nt, (uint32_t)((uint64_t)nt * eighths / 8), cap, cur, loc, rem);
            cudaEventRecord(b);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms, a, b);
        }
        printf("frontier scatter, %d/8 of the tiles in peer memory: %7.2f ms per 1e9 records  (%.0f GB/s of 16-B records over NVLink)\n", eighths, ms * sc,
               eighths ? (double)N * 16 * eighths / 8 / (ms * 1e-3) / 1e9 : 0.0);
    }
    {
        const int64_t n = (int64_t)nt * cap;
        float ms = 0;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(a);
            copy_kernel<<<148 * 16, 256>>>(loc, rem, n);
            cudaEventRecord(b);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms, a, b);
        }
        printf("coalesced copy kernel to the peer: %.0f GB/s\n", (double)n * 16 / (ms * 1e-3) / 1e9);
    }
    return 0;
}
