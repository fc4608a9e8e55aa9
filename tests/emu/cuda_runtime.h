// tests/emu/cuda_runtime.h — TEST INFRASTRUCTURE, not a product path.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile monte_b200/csrc/*.cu unchanged into
// tests/emu/_build/libmonte_gpu_emu.so: every __global__ function becomes a plain function, a launch
// runs the CTAs of the grid one after the other, and the threads of a CTA are cooperative fibers
// (ucontext) that switch at __syncthreads() and at the warp collectives (__shfl_sync, __ballot_sync,
// __reduce_add_sync ...).  "Device memory" is host memory, streams execute in issue order.
//
// Purpose: the CPU test-suite (`pytest -m "not gpu"`, tests/test_emu_*.py) drives the SAME kernel source
// and the SAME host-side C ABI code (argument checks, scratch buffers, chunking, launch geometry) as the
// GPU build and compares it with the oracle, so indexing, synchronisation and scheduling logic are
// checked without a GPU.  It proves nothing about speed, and fp32 results differ from the GPU's in the
// last ulp (nvcc contracts a*b+c into FMAs, g++ is told not to; MUFU approximations become libm calls).
// libmonte_gpu.so never contains this code, monte_b200.api never loads the emulation library, and there
// is no CPU fallback: without a B200 the product fails loudly (MONTE_E_NODEV).
#pragma once
#define MONTE_EMU 1

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

// ---- qualifiers (defined after the standard headers: libstdc++ spells some of these inside
// __attribute__((...)) itself) ----------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static          // CTAs run one at a time, so one static copy per kernel is per-CTA

// ---- vector types ---------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- the SIMT emulation ---------------------------------------------------------------------------
namespace monte_emu {

struct ThreadCtx {              // what a kernel sees through threadIdx / blockIdx / blockDim / gridDim
    uint3 tid, bid;
    dim3 bdim, gdim;
    int lin;                    // linear thread index in the CTA
};
ThreadCtx &cur();
void *dyn_smem();
void cta_barrier();                                         // __syncthreads
// every lane of the warp that is still running contributes one word; returns the mask of contributors
unsigned warp_gather(uint64_t mine, uint64_t out[32]);
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);

struct Cfg { dim3 grid, block; size_t smem; };
template <class... B> struct Bound { Cfg c; std::tuple<B...> args; };
struct CfgCall {
    Cfg c;
    template <class... B> Bound<std::decay_t<B>...> operator()(B &&...b) const {
        return Bound<std::decay_t<B>...>{c, std::tuple<std::decay_t<B>...>(std::forward<B>(b)...)};
    }
};
// `kernel MONTE_CFG(grid, block, smem, stream)(args...)` expands to `kernel * CfgCall{...}(args...)`
template <class... A, class... B> void operator*(void (*k)(A...), const Bound<B...> &b) {
    launch(b.c.grid, b.c.block, b.c.smem, [&] { std::apply([&](const auto &...x) { k(x...); }, b.args); });
}

template <class T> inline uint64_t to_word(T v) {
    static_assert(sizeof(T) <= 8, "warp collectives move at most 8 bytes");
    uint64_t w = 0;
    memcpy(&w, &v, sizeof(T));
    return w;
}
template <class T> inline T from_word(uint64_t w) {
    T v;
    memcpy(&v, &w, sizeof(T));
    return v;
}

}  // namespace monte_emu

#define threadIdx (::monte_emu::cur().tid)
#define blockIdx (::monte_emu::cur().bid)
#define blockDim (::monte_emu::cur().bdim)
#define gridDim (::monte_emu::cur().gdim)

static inline void __syncthreads() { ::monte_emu::cta_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { uint64_t o[32]; ::monte_emu::warp_gather(0, o); }

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    uint64_t o[32];
    ::monte_emu::warp_gather(::monte_emu::to_word(v), o);
    return ::monte_emu::from_word<T>(o[src & 31]);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lanemask, int = 32) {
    uint64_t o[32];
    ::monte_emu::warp_gather(::monte_emu::to_word(v), o);
    return ::monte_emu::from_word<T>(o[(::monte_emu::cur().lin ^ lanemask) & 31]);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    uint64_t o[32];
    const unsigned part = ::monte_emu::warp_gather(pred ? 1u : 0u, o);
    unsigned m = 0;
    for (int l = 0; l < 32; l++) if (o[l]) m |= 1u << l;
    return m & part;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    uint64_t o[32];
    const unsigned live = ::monte_emu::warp_gather(v, o);
    unsigned s = 0;
    for (int l = 0; l < 32; l++) if (live & (1u << l)) s += (unsigned)o[l];
    return s;
}

// ---- device intrinsics ------------------------------------------------------------------------------
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __uint_as_float(unsigned u) { return ::monte_emu::from_word<float>(u); }
static inline float __int_as_float(int u) { return ::monte_emu::from_word<float>((unsigned)u); }
static inline int __float_as_int(float f) { return (int)::monte_emu::to_word(f); }
static inline unsigned __float_as_uint(float f) { return (unsigned)::monte_emu::to_word(f); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
// glibc declares __logf/__expf/__sincosf itself (internal aliases of logf/expf/sincosf): use macros
#define __logf(x) logf(x)
#define __expf(x) expf(x)
#define __sincosf(x, s, c) sincosf((x), (s), (c))
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline void sincospi(double x, double *s, double *c) {
    // exact at the multiples of 1/2 like CUDA's sincospi
    double r = fmod(x, 2.0);
    if (r < 0) r += 2.0;
    if (r == 0.0) { *s = 0; *c = 1; }
    else if (r == 0.5) { *s = 1; *c = 0; }
    else if (r == 1.0) { *s = 0; *c = -1; }
    else if (r == 1.5) { *s = -1; *c = 0; }
    else { *s = sin(M_PI * r); *c = cos(M_PI * r); }
}
// the translation units are compiled with -ffp-contract=off, so these are the plain IEEE operations
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }

// fibers of one OS thread: a plain read-modify-write is atomic
template <class T, class U> static inline T atomicAdd(T *p, U v) { const T old = *p; *p = (T)(old + (T)v); return old; }

// CUDA's global min/max overload set (the subset the sources use)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }

// ---- runtime API (the subset libmonte_gpu uses) -------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorPeerAccessAlreadyEnabled = 704 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct CUstream_st;
typedef CUstream_st *cudaStream_t;
struct CUevent_st { double t_ms; };
typedef CUevent_st *cudaEvent_t;
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
};

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int dev);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaDeviceCanAccessPeer(int *can, int dev, int peer);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t monte_emu_malloc(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return monte_emu_malloc((void **)p, bytes); }
cudaError_t cudaFree(void *p);
enum { cudaHostAllocPortable = 1 };
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned) { return monte_emu_malloc((void **)p, bytes); }
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemsetAsync(void *dst, int value, size_t n, cudaStream_t s);
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind, cudaStream_t s);
template <class T> static inline cudaError_t cudaFuncSetAttribute(T, cudaFuncAttribute, int) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, T, int, size_t) {
    *n = 1;
    return cudaSuccess;
}
