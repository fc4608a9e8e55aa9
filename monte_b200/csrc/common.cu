// common.cu — library lifetime, error text, scratch allocator.
#include "common.cuh"

namespace monte {

static Context g_ctx;
static thread_local char g_err[1024] = "";

Context &ctx() { return g_ctx; }

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
    Context &c = g_ctx;
    if (bytes <= c.scratch_bytes[slot]) return c.scratch[slot];
    if (c.scratch[slot]) cudaFree(c.scratch[slot]);
    c.scratch[slot] = nullptr;
    c.scratch_bytes[slot] = 0;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    c.scratch[slot] = p;
    c.scratch_bytes[slot] = bytes;
    return p;
}

}  // namespace monte

using namespace monte;

extern "C" {

int monte_gpu_abi_version(void) { return MONTE_GPU_ABI_VERSION; }

const char *monte_gpu_last_error(void) { return g_err; }

int monte_gpu_init(int ndev, const int *ids) {
    Context &c = ctx();
    MONTE_ARG(ndev == 1, "monte_gpu_init: this ABI version binds one device per process (ndev=%d)", ndev);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device visible (%s); libmonte_gpu has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return MONTE_E_NODEV;
    }
    const int dev = ids ? ids[0] : 0;
    MONTE_ARG(dev >= 0 && dev < count, "monte_gpu_init: device %d out of range (0..%d)", dev, count - 1);
    cudaDeviceProp prop;
    MONTE_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
        return MONTE_E_NODEV;
    }
    if (c.inited && c.device == dev) return MONTE_OK;
    if (c.inited) monte_gpu_shutdown();
    MONTE_CUDA(cudaSetDevice(dev));
    MONTE_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    MONTE_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    c.device = dev;
    c.sm_count = prop.multiProcessorCount;
    c.inited = true;
    return MONTE_OK;
}

void monte_gpu_shutdown(void) {
    Context &c = ctx();
    if (!c.inited) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (auto f : g_cleanups) f();
    for (int i = 0; i < 12; i++) {
        if (c.scratch[i]) cudaFree(c.scratch[i]);
        c.scratch[i] = nullptr;
        c.scratch_bytes[i] = 0;
    }
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    c.stream = c.copy_stream = nullptr;
    c.inited = false;
}

int monte_gpu_sm_count(void) { return ctx().inited ? ctx().sm_count : 0; }

}  // extern "C"
