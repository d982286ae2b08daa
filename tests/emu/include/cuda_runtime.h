// CPU stand-in for the CUDA runtime + SIMT execution model (TEST INFRASTRUCTURE ONLY; tests/emu/README in build_emu.py).
// Every CUDA thread of a block is a fiber (ucontext) scheduled round-robin on one OS thread: __syncthreads() parks the
// fiber until all live threads of the block have arrived, the warp-synchronous intrinsics (__ballot_sync,
// __shfl_*_sync, __match_any_sync, ...) exchange through a per-warp buffer between two warp barriers, atomics are plain
// read-modify-writes, __shared__ variables are function-local statics (blocks run one after the other).  A thread that
// returns no longer counts for any barrier, as on the GPU; a barrier that can never complete (reached by only part of
// its threads) is reported as a deadlock instead of hanging.  ABK_EMU_SHUFFLE=<seed> randomises the order in which the
// fibers of a block are resumed, so code that silently relies on lane/warp execution order (a missing barrier) shows up
// as a wrong result.
// Kernel launches `k<<<grid, block, smem, stream>>>(args)` are rewritten to `emu::launch(grid, block, smem, stream)(k)(args)`
// by tests/emu/build_emu.py.  Only what libabk's ctx / ingest / kfields sources use is provided.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
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

// ---- SIMT execution: one fiber (ucontext) per CUDA thread, cooperative round-robin scheduling ----------------
namespace emu {
enum Wait { RUNNABLE = 0, AT_BLOCK_BARRIER, AT_WARP_BARRIER, DONE };
struct Fiber {
    ucontext_t uc;
    unsigned char *stack = nullptr;
    dim3 tid, bid;
    int lane = 0, warp = 0;
    Wait wait = RUNNABLE;
};
struct WarpState {
    uint64_t buf[32];
    int live = 0, waiting = 0;
};
struct BlockState {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    std::vector<unsigned char> dyn;
    int live = 0, waiting = 0;
};
inline BlockState *g_block = nullptr;
inline Fiber *g_cur = nullptr;
inline ucontext_t g_sched;
inline const std::function<void()> *g_body = nullptr;
inline void *dyn_smem() { return g_block->dyn.data(); }
inline void yield() { swapcontext(&g_cur->uc, &g_sched); }

inline void fiber_main()
{
    (*g_body)();
    Fiber *f = g_cur;
    f->wait = DONE;
    g_block->live--;
    g_block->warps[f->warp].live--;
    g_block->warps[f->warp].buf[f->lane] = 0;
    swapcontext(&f->uc, &g_sched);
}

// release the barriers whose live participants have all arrived (an exited thread no longer counts, as on the GPU)
inline void release_barriers(BlockState &b)
{
    if (b.live > 0 && b.waiting == b.live) {
        for (auto &f : b.fibers) if (f.wait == AT_BLOCK_BARRIER) f.wait = RUNNABLE;
        b.waiting = 0;
    }
    for (size_t w = 0; w < b.warps.size(); w++) {
        WarpState &ws = b.warps[w];
        if (ws.live > 0 && ws.waiting == ws.live) {
            for (auto &f : b.fibers) if (f.warp == (int)w && f.wait == AT_WARP_BARRIER) f.wait = RUNNABLE;
            ws.waiting = 0;
        }
    }
}

inline void block_barrier()
{
    g_cur->wait = AT_BLOCK_BARRIER;
    g_block->waiting++;
    yield();
}
inline void warp_barrier()
{
    g_cur->wait = AT_WARP_BARRIER;
    g_block->warps[g_cur->warp].waiting++;
    yield();
}

struct Cfg {
    dim3 g, b;
    size_t smem;
};
inline dim3 g_blockDim, g_gridDim;

inline void run(const Cfg &c, const std::function<void()> &body)
{
    g_blockDim = c.b;
    g_gridDim = c.g;
    g_body = &body;
    const int n = (int)(c.b.x * c.b.y * c.b.z);
    constexpr size_t STACK = 256 * 1024;
    static std::vector<unsigned char *> pool;   // fiber stacks, allocated once and never touched up front
    while ((int)pool.size() < n) pool.push_back((unsigned char *)std::malloc(STACK));
    BlockState blk;
    blk.fibers.resize(n);
    for (int t = 0; t < n; t++) blk.fibers[t].stack = pool[t];
    blk.warps.resize((n + 31) / 32);
    g_block = &blk;
    static const char *shuffle_env = std::getenv("ABK_EMU_SHUFFLE");
    const bool shuffle = shuffle_env != nullptr && shuffle_env[0] != 0;
    static unsigned long long rng = shuffle ? std::strtoull(shuffle_env, nullptr, 10) * 2654435761ull + 1 : 1;
    std::vector<int> order(n);
    for (int t = 0; t < n; t++) order[t] = t;
    for (unsigned bz = 0; bz < c.g.z; bz++)
        for (unsigned by = 0; by < c.g.y; by++)
            for (unsigned bx = 0; bx < c.g.x; bx++) {
                blk.dyn.assign(c.smem + 16, 0);
                blk.live = n;
                blk.waiting = 0;
                for (auto &w : blk.warps) { std::memset(w.buf, 0, sizeof(w.buf)); w.live = 0; w.waiting = 0; }
                for (int t = 0; t < n; t++) {
                    Fiber &f = blk.fibers[t];
                    f.tid = dim3(t % c.b.x, (t / c.b.x) % c.b.y, t / (c.b.x * c.b.y));
                    f.bid = dim3(bx, by, bz);
                    f.lane = t % 32;
                    f.warp = t / 32;
                    f.wait = RUNNABLE;
                    blk.warps[f.warp].live++;
                    getcontext(&f.uc);
                    f.uc.uc_stack.ss_sp = f.stack;
                    f.uc.uc_stack.ss_size = STACK;
                    f.uc.uc_link = nullptr;
                    makecontext(&f.uc, (void (*)())fiber_main, 0);
                }
                while (blk.live > 0) {
                    bool progressed = false;
                    if (shuffle) {   // ABK_EMU_SHUFFLE: visit the fibers in a fresh random order every round (race hunting)
                        for (int t = n - 1; t > 0; t--) {
                            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                            std::swap(order[t], order[(rng >> 33) % (unsigned)(t + 1)]);
                        }
                    }
                    for (int q = 0; q < n; q++) {
                        const int t = order[q];
                        Fiber &f = blk.fibers[t];
                        if (f.wait != RUNNABLE) continue;
                        g_cur = &f;
                        swapcontext(&g_sched, &f.uc);
                        progressed = true;
                    }
                    release_barriers(blk);
                    if (!progressed) {
                        bool any = false;
                        for (auto &f : blk.fibers) any |= (f.wait == RUNNABLE);
                        if (!any) {
                            std::fprintf(stderr, "emu: deadlock in block (%u,%u,%u): %d live threads, %d at the block barrier; "
                                         "a barrier or warp-synchronous intrinsic is reached by only part of its threads\n",
                                         bx, by, bz, blk.live, blk.waiting);
                            std::abort();
                        }
                    }
                }
            }
    g_block = nullptr;
    g_cur = nullptr;
}

template <class F>
struct Bound {
    Cfg c;
    F f;
    template <class... A>
    void operator()(A... a)
    {
        std::function<void()> body = [&] { f(a...); };
        run(c, body);
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
    WarpState &w = g_block->warps[g_cur->warp];
    const int lane = g_cur->lane;
    w.buf[lane] = to_bits(v);
    warp_barrier();
    const int s = src(lane);
    const T r = (s >= 0 && s < 32) ? from_bits<T>(w.buf[s]) : v;
    warp_barrier();
    return r;
}
}  // namespace emu

#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_cur->bid)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
inline unsigned __ballot_sync(unsigned, int pred)
{
    emu::WarpState &w = emu::g_block->warps[emu::g_cur->warp];
    const int lane = emu::g_cur->lane;
    w.buf[lane] = pred ? 1 : 0;
    emu::warp_barrier();
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r |= (unsigned)(w.buf[l] != 0) << l;
    emu::warp_barrier();
    w.buf[lane] = 0;
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
template <class T>
inline unsigned __match_any_sync(unsigned, T v)
{
    emu::WarpState &w = emu::g_block->warps[emu::g_cur->warp];
    const int lane = emu::g_cur->lane;
    w.buf[lane] = emu::to_bits(v) ^ 0x8000000000000000ull;   // exited lanes hold 0: never equal to a live value
    emu::warp_barrier();
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r |= (unsigned)(w.buf[l] == w.buf[lane]) << l;
    emu::warp_barrier();
    w.buf[lane] = 0;
    return r;
}

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
using std::max;
using std::min;
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
inline long long min(long long a, int b) { return a < b ? a : b; }
inline long long max(long long a, int b) { return a > b ? a : b; }
inline void sincospif(float x, float *s, float *c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }
inline void sincospi(double x, double *s, double *c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
template <class T>
inline T __ldg(const T *p) { return *p; }
template <class T>
inline T __ldcs(const T *p) { return *p; }

template <class T>
inline T atomicAdd(T *p, T v) { const T old = *p; *p = old + v; return old; }
inline unsigned atomicAdd(unsigned *p, int v) { return atomicAdd<unsigned>(p, (unsigned)v); }
template <class T>
inline T atomicSub(T *p, T v) { const T old = *p; *p = old - v; return old; }
template <class T>
inline T atomicExch(T *p, T v) { const T old = *p; *p = v; return old; }

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
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
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
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
struct cudaFuncAttributes { size_t sharedSizeBytes = 4096; int numRegs = 0; };
template <class F>
inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, F) { *a = cudaFuncAttributes(); return cudaSuccess; }

// ---- vector types ----------------------------------------------------------------------------------
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct double2 { double x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
// packed float32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): two independent IEEE operations
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
