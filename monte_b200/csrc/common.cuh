// common.cuh — process-wide context and error plumbing of libmonte_gpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/monte_gpu.h"

// Kernel launches and dynamic shared memory are spelled through these two macros so that the same
// sources also compile as plain C++ against tests/emu/cuda_runtime.h, a SIMT emulation the CPU test-suite
// uses to drive the kernels and this host code without a GPU.  MONTE_EMU is never defined in the product
// build (monte_b200/build.py, nvcc only); there is no CPU fallback.
#ifdef MONTE_EMU
#define MONTE_CFG(grid, block, smem, stream) \
    *::monte_emu::CfgCall{::monte_emu::Cfg{dim3(grid), dim3(block), (size_t)(smem)}}
#define MONTE_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(::monte_emu::dyn_smem())
#else
#define MONTE_CFG(grid, block, smem, stream) <<<(grid), (block), (smem), (stream)>>>
#define MONTE_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

namespace monte {

constexpr int MAX_DEV = 8;         // devices one process may bind (monte_gpu_init(ndev, ids)): one NVSwitch box
constexpr int N_SCRATCH = 16;

// One Context per bound device.  Every entry point works on the *current* device (index cur_dev(), 0 after
// monte_gpu_init); the multi-device orchestration (multi.cu) walks the devices with use_dev(i), issues the same
// single-device code on each of them, and returns to device 0.
struct Context {
    bool      inited = false;
    int       device = 0;              // CUDA ordinal
    int       sm_count = 0;
    cudaStream_t stream = nullptr;     // library-owned stream for the host-buffer entry points
    cudaStream_t copy_stream = nullptr; // second stream: D2H of finished slabs overlaps compute
    cudaStream_t aux_stream = nullptr;  // third stream: multi-device FDK backprojects on it while `stream` still filters
    // grow-only device scratch shared by the host-buffer entry points
    void  *scratch[N_SCRATCH] = {nullptr};
    size_t scratch_bytes[N_SCRATCH] = {0};
};

Context &ctx();                    // of the current device
Context &ctx_of(int i);
int  n_dev();                      // devices bound by monte_gpu_init (0 before)
int  cur_dev();                    // index of the current device in the bound list
int  use_dev(int i);               // cudaSetDevice + make it current; MONTE_OK or an error code
bool peers_ok();                   // every bound device can load/store every other one's memory (NVLink P2P enabled)
// per-device module state (function attributes already raised, cached tables, events ...): one copy per bound device
template <class T> struct PerDev {
    T v[MAX_DEV];
    T &get() { return v[cur_dev()]; }
    T &of(int i) { return v[i]; }
};
// modules register a function that frees their cached device buffers (called by monte_gpu_shutdown)
void at_shutdown(void (*fn)());
void set_error(const char *fmt, ...);
int  cuda_fail(cudaError_t e, const char *what, const char *file, int line);
// returns device pointer of at least `bytes` in slot `slot` (contents undefined), or nullptr
void *scratch(int slot, size_t bytes);

#define MONTE_CUDA(call)                                                        \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return ::monte::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define MONTE_REQUIRE_INIT()                                                    \
    do {                                                                        \
        if (!::monte::ctx().inited) {                                           \
            ::monte::set_error("monte_gpu_init has not been called");           \
            return MONTE_E_NOINIT;                                              \
        }                                                                       \
    } while (0)

#define MONTE_ARG(cond, ...)                                                    \
    do {                                                                        \
        if (!(cond)) {                                                          \
            ::monte::set_error(__VA_ARGS__);                                    \
            return MONTE_E_ARG;                                                 \
        }                                                                       \
    } while (0)

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    explicit EventTimer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~EventTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, s); }
    void stop() { cudaEventRecord(b, s); }
    double ms() { float m = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&m, a, b); return m; }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace monte
