// common.cu — library lifetime (one process, 1..8 devices), error text, scratch allocator, NCCL loader.
#include "common.cuh"
#include "nccl_dl.cuh"
#ifndef MONTE_EMU
#include <dlfcn.h>
#endif

namespace monte {

static Context g_ctxs[MAX_DEV];
static int g_ndev = 0;
// (thread-local: a multi-device call may issue each device's launches from a thread of its own, see fdk_multi)
static thread_local int g_cur = 0;
static bool g_peers = false;
static thread_local char g_err[1024] = "";

Context &ctx() { return g_ctxs[g_cur]; }
Context &ctx_of(int i) { return g_ctxs[i]; }
int n_dev() { return g_ndev; }
int cur_dev() { return g_cur; }
bool peers_ok() { return g_peers; }

int use_dev(int i) {
    if (i < 0 || i >= g_ndev) { set_error("use_dev: device index %d out of range (bound: %d)", i, g_ndev); return MONTE_E_ARG; }
    MONTE_CUDA(cudaSetDevice(g_ctxs[i].device));
    g_cur = i;
    return MONTE_OK;
}

static std::vector<void (*)()> g_cleanups;
void at_shutdown(void (*fn)()) {
    for (auto f : g_cleanups) if (f == fn) return;
    g_cleanups.push_back(fn);
}

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return MONTE_E_CUDA;
}

void *scratch(int slot, size_t bytes) {
    Context &c = ctx();
    if (bytes <= c.scratch_bytes[slot]) return c.scratch[slot];
    if (c.scratch[slot]) cudaFree(c.scratch[slot]);
    c.scratch[slot] = nullptr;
    c.scratch_bytes[slot] = 0;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %zu bytes failed on device %d: %s", bytes, c.device, cudaGetErrorString(e));
        return nullptr;
    }
    c.scratch[slot] = p;
    c.scratch_bytes[slot] = bytes;
    return p;
}

// ---- NCCL, loaded at run time ------------------------------------------------------------------------
// libmonte_gpu.so has no link-time dependency on NCCL: a single-device process never touches it, and a host that
// already carries an NCCL (PyTorch bundles its own libnccl.so.2) keeps exactly one copy in the process.
static NcclApi g_nccl;
static ncclComm_t g_comms[MAX_DEV] = {nullptr};
static bool g_comms_ok = false;

#ifdef MONTE_EMU
bool nccl_emu_fill(NcclApi &a);             // tests/emu/emu_runtime.cpp: an in-process stand-in (tests only)
#endif

const NcclApi *nccl_api() {
    if (g_nccl.loaded) return &g_nccl;
#ifdef MONTE_EMU
    nccl_emu_fill(g_nccl);
    g_nccl.loaded = true;
    return &g_nccl;
#else
    void *h = nullptr;
    const char *names[] = {getenv("MONTE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { set_error("NCCL not found (dlopen libnccl.so.2: %s); set MONTE_NCCL_LIB", dlerror()); return nullptr; }
#define MONTE_NCCL_SYM(field, name)                                                    \
    do {                                                                               \
        *(void **)(&g_nccl.field) = dlsym(h, name);                                    \
        if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); return nullptr; } \
    } while (0)
    MONTE_NCCL_SYM(CommInitAll, "ncclCommInitAll");
    MONTE_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    MONTE_NCCL_SYM(GetErrorString, "ncclGetErrorString");
    MONTE_NCCL_SYM(GroupStart, "ncclGroupStart");
    MONTE_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    MONTE_NCCL_SYM(Reduce, "ncclReduce");
    MONTE_NCCL_SYM(Send, "ncclSend");
    MONTE_NCCL_SYM(Recv, "ncclRecv");
    MONTE_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef MONTE_NCCL_SYM
    g_nccl.loaded = true;
    return &g_nccl;
#endif
}

// one communicator per bound device (ncclCommInitAll), created on first use
int nccl_comms(ncclComm_t **out) {
    if (!g_comms_ok) {
        const NcclApi *n = nccl_api();
        if (!n) return MONTE_E_CUDA;
        int ids[MAX_DEV];
        for (int i = 0; i < g_ndev; i++) ids[i] = g_ctxs[i].device;
        const int rc = n->CommInitAll(g_comms, g_ndev, ids);
        if (rc != 0) { set_error("ncclCommInitAll over %d devices failed: %s", g_ndev, n->GetErrorString(rc)); return MONTE_E_CUDA; }
        g_comms_ok = true;
        if (int rc2 = use_dev(g_cur)) return rc2;          // CommInitAll may leave another device current
    }
    *out = g_comms;
    return MONTE_OK;
}

}  // namespace monte

using namespace monte;

extern "C" {

int monte_gpu_abi_version(void) { return MONTE_GPU_ABI_VERSION; }

const char *monte_gpu_last_error(void) { return g_err; }

int monte_gpu_init(int ndev, const int *ids) {
    MONTE_ARG(ndev >= 1 && ndev <= MAX_DEV, "monte_gpu_init: ndev must be 1..%d (got %d)", MAX_DEV, ndev);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device visible (%s); libmonte_gpu has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return MONTE_E_NODEV;
    }
    int dev[MAX_DEV];
    for (int i = 0; i < ndev; i++) {
        dev[i] = ids ? ids[i] : i;
        MONTE_ARG(dev[i] >= 0 && dev[i] < count, "monte_gpu_init: device %d out of range (0..%d)", dev[i], count - 1);
        for (int j = 0; j < i; j++) MONTE_ARG(dev[j] != dev[i], "monte_gpu_init: device %d listed twice", dev[i]);
        cudaDeviceProp prop;
        MONTE_CUDA(cudaGetDeviceProperties(&prop, dev[i]));
        if (prop.major != 10) {
            set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev[i], prop.major, prop.minor);
            return MONTE_E_NODEV;
        }
    }
    if (g_ndev == ndev) {                         // same binding as before: nothing to do
        bool same = true;
        for (int i = 0; i < ndev; i++) same = same && g_ctxs[i].inited && g_ctxs[i].device == dev[i];
        if (same) return use_dev(0);
    }
    if (g_ndev) monte_gpu_shutdown();
    for (int i = 0; i < ndev; i++) {
        Context &c = g_ctxs[i];
        cudaDeviceProp prop;
        MONTE_CUDA(cudaGetDeviceProperties(&prop, dev[i]));
        MONTE_CUDA(cudaSetDevice(dev[i]));
        MONTE_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        MONTE_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        MONTE_CUDA(cudaStreamCreateWithFlags(&c.aux_stream, cudaStreamNonBlocking));
        c.device = dev[i];
        c.sm_count = prop.multiProcessorCount;
        c.inited = true;
    }
    g_ndev = ndev;
    // NVLink peer access between every pair: the multi-device paths load and store peer memory from kernels
    g_peers = ndev > 1;
    for (int i = 0; i < ndev && ndev > 1; i++) {
        MONTE_CUDA(cudaSetDevice(dev[i]));
        for (int j = 0; j < ndev; j++) {
            if (i == j) continue;
            int can = 0;
            MONTE_CUDA(cudaDeviceCanAccessPeer(&can, dev[i], dev[j]));
            if (!can) { g_peers = false; continue; }
            e = cudaDeviceEnablePeerAccess(dev[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();        // (another library of this process did it)
            else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
        }
    }
    return use_dev(0);
}

void monte_gpu_shutdown(void) {
    if (!g_ndev) return;
    for (int i = 0; i < g_ndev; i++) { cudaSetDevice(g_ctxs[i].device); cudaDeviceSynchronize(); }
    if (g_comms_ok) {
        for (int i = 0; i < g_ndev; i++) if (g_comms[i]) { g_nccl.CommDestroy(g_comms[i]); g_comms[i] = nullptr; }
        g_comms_ok = false;
    }
    for (int i = 0; i < g_ndev; i++) {           // module caches are per device: every cleanup runs on every device
        g_cur = i;
        cudaSetDevice(g_ctxs[i].device);
        for (auto f : g_cleanups) f();
    }
    for (int d = 0; d < g_ndev; d++) {
        Context &c = g_ctxs[d];
        cudaSetDevice(c.device);
        for (int i = 0; i < N_SCRATCH; i++) {
            if (c.scratch[i]) cudaFree(c.scratch[i]);
            c.scratch[i] = nullptr;
            c.scratch_bytes[i] = 0;
        }
        if (c.stream) cudaStreamDestroy(c.stream);
        if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
        if (c.aux_stream) cudaStreamDestroy(c.aux_stream);
        c.stream = c.copy_stream = c.aux_stream = nullptr;
        c.inited = false;
    }
    g_ndev = 0; g_cur = 0; g_peers = false;
}

int monte_gpu_sm_count(void) { return ctx().inited ? ctx().sm_count : 0; }
int monte_gpu_device_count(void) { return g_ndev; }
int monte_gpu_peer_access(void) { return g_peers ? 1 : 0; }

}  // extern "C"
