// Micro-benchmark 8 (round 2): does the row padding of the in-place R2C layout matter to cuFFT?  1024^3, complex row length
// 513 (the standard 2(n/2+1) floats), 520, 528, 544, 576 -- a row of 513 complex numbers starts 8 bytes off a 64-byte boundary
// in every other row, which the strided y- and x-passes might not like.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fft_pad scripts/micro/fft_pad_micro.cu -lcufft && /tmp/fft_pad
#include <cstdio>
#include <cuda_runtime.h>
#include <cufft.h>

__global__ void fill_kernel(float *g, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)i * 2654435761u; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        g[i] = (x >> 8) * (1.0f / 16777216.0f) - 0.5f;
    }
}

int main()
{
    const long long n = 1024;
    const long long pads[] = {513, 516, 520, 528, 544, 576};
    for (long long nzc : pads) {
        float *grid; void *work;
        const size_t bytes = (size_t)n * n * nzc * 8;
        if (cudaMalloc(&grid, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
        cudaMemset(grid, 0, bytes);
        fill_kernel<<<1184, 256>>>(grid, (long long)(bytes / 4));      // pseudo-random data (zeros could flatter the memory system)
        cufftHandle h;
        cufftCreate(&h);
        cufftSetAutoAllocation(h, 0);
        long long dims[3] = {n, n, n}, rembed[3] = {n, n, 2 * nzc}, cembed[3] = {n, n, nzc};
        size_t w = 0;
        cufftResult r = cufftMakePlanMany64(h, 3, dims, rembed, 1, n * n * 2 * nzc, cembed, 1, n * n * nzc, CUFFT_R2C, 1, &w);
        if (r != CUFFT_SUCCESS) { printf("nzc %lld: plan error %d\n", nzc, (int)r); continue; }
        cudaMalloc(&work, w ? w : 16);
        cufftSetWorkArea(h, work);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        float tot = 0;
        for (int i = 0; i < 4; i++) {
            fill_kernel<<<1184, 256>>>(grid, (long long)(bytes / 4));
            cudaEventRecord(a);
            cufftExecR2C(h, grid, (cufftComplex *)grid);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (i) tot += ms;
        }
        printf("complex row length %lld: in-place 3-D R2C %.2f ms  (work area %zu bytes)\n", nzc, tot / 3, w);
        cufftDestroy(h);
        cudaFree(work); cudaFree(grid);
    }
    return 0;
}
