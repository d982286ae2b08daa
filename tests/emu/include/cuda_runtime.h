// CPU stand-in for the CUDA runtime + SIMT execution model (TEST INFRASTRUCTURE ONLY; tests/emu/README in build_emu.py).
// Every CUDA thread of a block is a real OS thread: __syncthreads() is a std::barrier over the block, the
// warp-synchronous intrinsics (__ballot_sync, __shfl_*_sync, ...) exchange through a per-warp buffer guarded by a
// std::barrier over the warp's lanes, atomics are real atomics, __shared__ variables are function-local statics
// (blocks run one after the other).  A thread that returns drops out of both barriers, as on the GPU.
// Kernel launches `k<<<grid, block, smem, stream>>>(args)` are rewritten to `emu::launch(grid, block, smem, stream)(k)(args)`
// by tests/emu/build_emu.py.  Only what libabk's ctx / ingest / kfields sources use is provided.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned long long x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x((unsigned)x_), y(y_), z(z_) {}
};

namespace emu {
struct Warp {
    std::barrier<> bar;
    uint64_t buf[32];
    explicit Warp(int lanes) : bar(lanes) { std::memset(buf, 0, sizeof(buf)); }
};
struct Block {
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    std::vector<unsigned char> dyn;
    explicit Block(int n) : bar(n) {}
};
inline thread_local Block *t_block = nullptr;
inline thread_local Warp *t_warp = nullptr;
inline thread_local int t_lane = 0;
inline void *dyn_smem() { return t_block->dyn.data(); }
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace emu {
struct Cfg {
    dim3 g, b;
    size_t smem;
};
template <class Body>
void run(const Cfg &c, Body body)
{
    blockDim = c.b;
    gridDim = c.g;
    const int nthreads = (int)(c.b.x * c.b.y * c.b.z);
    for (unsigned bz = 0; bz < c.g.z; bz++)
        for (unsigned by = 0; by < c.g.y; by++)
            for (unsigned bx = 0; bx < c.g.x; bx++) {
                Block blk(nthreads);
                blk.dyn.assign(c.smem + 16, 0);
                for (int w = 0; w * 32 < nthreads; w++) blk.warps.emplace_back(new Warp(std::min(32, nthreads - w * 32)));
                std::vector<std::thread> th;
                th.reserve(nthreads);
                for (int t = 0; t < nthreads; t++)
                    th.emplace_back([&, t] {
                        threadIdx = dim3(t % c.b.x, (t / c.b.x) % c.b.y, t / (c.b.x * c.b.y));
                        blockIdx = dim3(bx, by, bz);
                        t_block = &blk;
                        t_warp = blk.warps[t / 32].get();
                        t_lane = t % 32;
                        body();
                        t_warp->buf[t_lane] = 0;
                        t_warp->bar.arrive_and_drop();
                        blk.bar.arrive_and_drop();
                    });
                for (auto &x : th) x.join();
            }
}
template <class F>
struct Bound {
    Cfg c;
    F f;
    template <class... A>
    void operator()(A... a)
    {
        run(c, [&] { f(a...); });
    }
};
struct Launch {
    Cfg c;
    template <class F>
    Bound<F> operator()(F f) { return Bound<F>{c, f}; }
};
inline Launch launch(dim3 g, dim3 b, size_t smem = 0, void * = nullptr) { return Launch{Cfg{g, b, smem}}; }

template <class T>
inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "shuffle operand too wide");
    uint64_t u = 0;
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T>
inline T from_bits(uint64_t u)
{
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}
// every lane publishes a value, then reads the lane picked by `src(lane)` (its own value if out of range)
template <class T, class Pick>
inline T exchange(T v, Pick src)
{
    Warp *w = t_warp;
    w->buf[t_lane] = to_bits(v);
    w->bar.arrive_and_wait();
    const int s = src(t_lane);
    const T r = (s >= 0 && s < 32) ? from_bits<T>(w->buf[s]) : v;
    w->bar.arrive_and_wait();
    return r;
}
}  // namespace emu

inline void __syncthreads() { emu::t_block->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_warp->bar.arrive_and_wait(); }
inline unsigned __ballot_sync(unsigned, int pred)
{
    emu::Warp *w = emu::t_warp;
    w->buf[emu::t_lane] = pred ? 1 : 0;
    w->bar.arrive_and_wait();
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r |= (unsigned)(w->buf[l] != 0) << l;
    w->bar.arrive_and_wait();
    w->buf[emu::t_lane] = 0;
    return r;
}
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::exchange(v, [=](int) { return src & 31; }); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return emu::exchange(v, [=](int l) { return l - (int)d; }); }
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return emu::exchange(v, [=](int l) { return l + (int)d; }); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu::exchange(v, [=](int l) { return l ^ m; }); }

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline double __longlong_as_double(long long v) { return emu::from_bits<double>((uint64_t)v); }
inline float __uint_as_float(unsigned v) { return emu::from_bits<float>(v); }
inline unsigned __float_as_uint(float v) { return (unsigned)emu::to_bits(v); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
template <class T>
inline T __ldg(const T *p) { return *p; }
template <class T>
inline T __ldcs(const T *p) { return *p; }

template <class T>
inline T atomicAdd(T *p, T v)
{
    if constexpr (std::is_integral_v<T>) {
        return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
    } else {
        std::atomic_ref<T> a(*p);
        T old = a.load(std::memory_order_relaxed);
        while (!a.compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {}
        return old;
    }
}
inline unsigned atomicAdd(unsigned *p, int v) { return atomicAdd<unsigned>(p, (unsigned)v); }
template <class T>
inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }

// ---- runtime API ------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp {
    int major, minor, multiProcessorCount;
    size_t sharedMemPerBlockOptin;
};
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp{10, 0, 4, 227 * 1024}; return cudaSuccess; }
template <class T>
inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)std::malloc(n); return cudaSuccess; }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- vector types ----------------------------------------------------------------------------------
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct double2 { double x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
