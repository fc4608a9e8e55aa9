// tests/emu/emu_runtime.cpp — TEST INFRASTRUCTURE: the fiber scheduler and the fake runtime behind
// tests/emu/cuda_runtime.h (see the header for what this is and is not).
#include "cuda_runtime.h"

#include <chrono>
#include <sys/mman.h>
#include <ucontext.h>

namespace monte_emu {

namespace {

constexpr size_t STACK_BYTES = 256 * 1024;
constexpr size_t SMEM_BYTES = 227 * 1024;       // the opt-in maximum of dynamic shared memory per CTA on sm_100

struct Barrier {
    int arrived = 0;
    unsigned gen = 0;
    int live = 0;
};
struct Warp {
    Barrier bar;
    uint64_t buf[2][32];
    unsigned mask[2] = {0, 0};            // lanes that contributed to the collective held in buf[slot]
    unsigned mask_gen[2] = {0, 0};
};
struct Fiber {
    ucontext_t uc;
    ThreadCtx tc;
    bool done = false;
    char *stack = nullptr;
};

struct Machine {
    std::vector<char *> stacks;           // reused between CTAs
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    Barrier cta;
    ucontext_t sched;
    Fiber *running = nullptr;
    const std::function<void()> *body = nullptr;
    bool progress = false;
    char *smem = nullptr;                 // dynamic shared memory of the running launch: a heap block of exactly the
    size_t smem_bytes = 0;                // requested size, so the ASan build reports overruns of it as well
};
Machine *g_m = nullptr;
ThreadCtx g_host_ctx;                     // cur() outside a launch

char *get_stack(Machine &m, size_t i) {
    while (m.stacks.size() <= i) {
        void *p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("monte_emu: mmap of a fiber stack"); abort(); }
        m.stacks.push_back((char *)p);
    }
    return m.stacks[i];
}

void yield() {
    Machine &m = *g_m;
    Fiber *f = m.running;
    swapcontext(&f->uc, &m.sched);
}

// a barrier over `live` participants; exited threads no longer count (CUDA semantics since sm_70)
void release_if_complete(Barrier &b) {
    if (b.live > 0 && b.arrived >= b.live) { b.arrived = 0; b.gen++; g_m->progress = true; }
}
void wait(Barrier &b) {
    const unsigned g = b.gen;
    b.arrived++;
    release_if_complete(b);
    while (b.gen == g) yield();
}

void fiber_main() {
    Machine &m = *g_m;
    Fiber *f = m.running;
    (*m.body)();
    f->done = true;
    m.progress = true;
    Warp &w = m.warps[f->tc.lin >> 5];        // (its last contributions stay readable for the lanes still running)
    w.bar.live--; m.cta.live--;
    release_if_complete(w.bar);
    release_if_complete(m.cta);
    swapcontext(&f->uc, &m.sched);          // never resumed
}

}  // namespace

ThreadCtx &cur() { return g_m && g_m->running ? g_m->running->tc : g_host_ctx; }
void *dyn_smem() { return g_m->smem; }
void cta_barrier() { wait(g_m->cta); }
unsigned warp_gather(uint64_t mine, uint64_t out[32]) {
    Machine &m = *g_m;
    const int lin = m.running->tc.lin;
    Warp &w = m.warps[lin >> 5];
    // consecutive collectives alternate buffers: a lane that runs ahead into collective n+1 cannot overwrite
    // words its neighbours have not read yet, and nobody reaches n+2 before everybody has arrived at n+1
    const int slot = w.bar.gen & 1;
    if (w.mask_gen[slot] != w.bar.gen + 1) {     // first arrival of this generation
        w.mask_gen[slot] = w.bar.gen + 1;
        w.mask[slot] = 0;
        memset(w.buf[slot], 0, sizeof(w.buf[slot]));
    }
    w.mask[slot] |= 1u << (lin & 31);
    w.buf[slot][lin & 31] = mine;
    wait(w.bar);
    memcpy(out, w.buf[slot], sizeof(w.buf[slot]));
    return w.mask[slot];
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    if (!g_m) g_m = new Machine();
    Machine &m = *g_m;
    if (m.running) { fprintf(stderr, "monte_emu: nested launch\n"); abort(); }
    if (smem > SMEM_BYTES) { fprintf(stderr, "monte_emu: %zu bytes of dynamic shared memory\n", smem); abort(); }
    const size_t nthr = (size_t)block.x * block.y * block.z;
    if (nthr == 0 || nthr > 1024) { fprintf(stderr, "monte_emu: bad block size %zu\n", nthr); abort(); }
    m.body = &body;
    free(m.smem);
    m.smem = nullptr; m.smem_bytes = smem;
    if (smem && posix_memalign((void **)&m.smem, 128, smem) != 0) { fprintf(stderr, "monte_emu: out of memory\n"); abort(); }
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                m.fibers.assign(nthr, Fiber());
                m.warps.assign((nthr + 31) / 32, Warp());
                m.cta = Barrier();
                m.cta.live = (int)nthr;
                if (smem) memset(m.smem, 0xFF, smem);        // uninitialised shared memory is garbage on the GPU too: NaNs here
                size_t i = 0;
                for (unsigned tz = 0; tz < block.z; tz++)
                    for (unsigned ty = 0; ty < block.y; ty++)
                        for (unsigned tx = 0; tx < block.x; tx++, i++) {
                            Fiber &f = m.fibers[i];
                            f.tc.tid = uint3{tx, ty, tz};
                            f.tc.bid = uint3{bx, by, bz};
                            f.tc.bdim = block; f.tc.gdim = grid; f.tc.lin = (int)i;
                            f.stack = get_stack(m, i);
                            Warp &w = m.warps[i >> 5];
                            w.bar.live++;
                            getcontext(&f.uc);
                            f.uc.uc_stack.ss_sp = f.stack;
                            f.uc.uc_stack.ss_size = STACK_BYTES;
                            f.uc.uc_link = &m.sched;
                            makecontext(&f.uc, (void (*)())fiber_main, 0);
                        }
                // MONTE_EMU_SCHED=reverse | random: the order in which runnable fibers are resumed.  A kernel whose
                // barriers are complete gives the same result for every order; a missing barrier does not.
                static const int sched = [] { const char *e = getenv("MONTE_EMU_SCHED");
                                              return !e ? 0 : !strcmp(e, "reverse") ? 1 : !strcmp(e, "random") ? 2 : 0; }();
                static uint64_t rng = 0x2545F4914F6CDD1Dull;
                std::vector<size_t> order(nthr);
                for (size_t k = 0; k < nthr; k++) order[k] = sched == 1 ? nthr - 1 - k : k;
                size_t left = nthr;
                while (left) {
                    m.progress = false;
                    if (sched == 2)
                        for (size_t k = nthr - 1; k > 0; k--) {         // Fisher-Yates with xorshift64*
                            rng ^= rng >> 12; rng ^= rng << 25; rng ^= rng >> 27;
                            std::swap(order[k], order[(size_t)((rng * 0x2545F4914F6CDD1Dull) >> 33) % (k + 1)]);
                        }
                    for (size_t kk = 0; kk < nthr; kk++) {
                        const size_t k = order[kk];
                        Fiber &f = m.fibers[k];
                        if (f.done) continue;
                        m.running = &f;
                        swapcontext(&m.sched, &f.uc);
                        m.running = nullptr;
                        if (f.done) left--;
                    }
                    if (left && !m.progress) {
                        fprintf(stderr, "monte_emu: deadlock in CTA (%u,%u,%u): %zu threads wait at barriers that cannot complete\n",
                                bx, by, bz, left);
                        abort();
                    }
                }
            }
    m.body = nullptr;
}

}  // namespace monte_emu

// ---- fake runtime -----------------------------------------------------------------------------------
static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "error"; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
// MONTE_EMU_DEVICES=n: n "devices" that all are this host's memory -- enough to drive the multi-device orchestration
// (per-device contexts, photon / view / z-slab partition, peer loads, NCCL call pattern) on the CPU
static int emu_device_count() { const char *e = getenv("MONTE_EMU_DEVICES"); const int n = e ? atoi(e) : 1; return n < 1 ? 1 : (n > 8 ? 8 : n); }
static int g_emu_dev = 0;
cudaError_t cudaGetDeviceCount(int *n) { *n = emu_device_count(); return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { const char *e = getenv("MONTE_EMU_NO_PEER"); *can = e && atoi(e) ? 0 : 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "monte_emu (CPU SIMT emulation, tests only)");
    p->major = 10; p->minor = 0;
    const char *e = getenv("MONTE_EMU_SMS");
    p->multiProcessorCount = e ? atoi(e) : 2;
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= emu_device_count()) return cudaErrorInvalidValue; g_emu_dev = d; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new CUevent_st{0.0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t_ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
cudaError_t monte_emu_malloc(void **p, size_t bytes) {
    // 256-byte alignment like cudaMalloc; freshly allocated device memory is not zero: poison it
    void *q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 1) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
    memset(q, 0xFF, bytes);     // (all-ones = NaN for float and double, -1 for integers)
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *dst, int value, size_t n, cudaStream_t) { memset(dst, value, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; r++) memmove((char *)dst + r * dpitch, (const char *)src + r * spitch, width);
    return cudaSuccess;
}

// ---- in-process stand-in for the NCCL entry points of csrc/nccl_dl.cuh (tests only) ----------------------------
// One host thread drives all "devices", exactly like the product's single-process multi-device mode: collective calls
// are queued between GroupStart and GroupEnd and executed at GroupEnd.  Reduce: the root's recv buffer receives the
// sum of all send buffers; Send/Recv: matched in issue order per (source, destination) pair.
#include "../../monte_b200/csrc/nccl_dl.cuh"
struct ncclComm { int rank, size; };
namespace {
struct EmuOp { int kind; const void *send; void *recv; size_t count; int dtype, root_or_peer, rank; };
std::vector<EmuOp> g_ops;
int g_group = 0;
size_t emu_dtype_size(int t) { return t == monte::NCCL_UINT8 ? 1 : t == monte::NCCL_INT64 ? 8 : 4; }
int emu_run_ops() {
    // reduces (kind 0): group by root
    std::vector<bool> done(g_ops.size(), false);
    for (size_t i = 0; i < g_ops.size(); i++) {
        if (done[i] || g_ops[i].kind != 0) continue;
        const EmuOp &a = g_ops[i];
        std::vector<const void *> sends;
        void *root_recv = nullptr;
        for (size_t j = i; j < g_ops.size(); j++) {
            if (done[j] || g_ops[j].kind != 0 || g_ops[j].root_or_peer != a.root_or_peer || g_ops[j].count != a.count) continue;
            sends.push_back(g_ops[j].send);
            if (g_ops[j].rank == a.root_or_peer) root_recv = g_ops[j].recv;
            done[j] = true;
        }
        if (!root_recv) return 3;
        if (a.dtype == monte::NCCL_INT32) {
            std::vector<int32_t> acc(a.count, 0);
            for (const void *s : sends) for (size_t k = 0; k < a.count; k++) acc[k] += ((const int32_t *)s)[k];
            memcpy(root_recv, acc.data(), a.count * 4);
        } else if (a.dtype == monte::NCCL_FLOAT32) {
            std::vector<float> acc(a.count, 0.f);
            for (const void *s : sends) for (size_t k = 0; k < a.count; k++) acc[k] += ((const float *)s)[k];
            memcpy(root_recv, acc.data(), a.count * 4);
        } else return 4;
    }
    // sends (kind 1) matched with recvs (kind 2) in issue order
    for (size_t i = 0; i < g_ops.size(); i++) {
        if (done[i] || g_ops[i].kind != 1) continue;
        bool found = false;
        for (size_t j = 0; j < g_ops.size() && !found; j++) {
            if (done[j] || g_ops[j].kind != 2) continue;
            if (g_ops[j].rank == g_ops[i].root_or_peer && g_ops[j].root_or_peer == g_ops[i].rank) {
                if (g_ops[j].count != g_ops[i].count) return 5;
                memmove(g_ops[j].recv, g_ops[i].send, g_ops[i].count * emu_dtype_size(g_ops[i].dtype));
                done[i] = done[j] = true; found = true;
            }
        }
        if (!found) return 6;
    }
    for (size_t i = 0; i < g_ops.size(); i++) if (!done[i]) return 7;
    g_ops.clear();
    return 0;
}
int emu_push(EmuOp op) { g_ops.push_back(op); return g_group ? 0 : emu_run_ops(); }
}  // namespace
namespace monte {
bool nccl_emu_fill(NcclApi &a) {
    a.CommInitAll = [](ncclComm_t *c, int n, const int *) { for (int i = 0; i < n; i++) c[i] = new ncclComm{i, n}; return 0; };
    a.CommDestroy = [](ncclComm_t c) { delete c; return 0; };
    a.GetErrorString = [](int r) -> const char * { return r == 0 ? "no error" : "monte_emu nccl: unmatched or malformed collective"; };
    a.GroupStart = []() { g_group++; return 0; };
    a.GroupEnd = []() { return --g_group == 0 ? emu_run_ops() : 0; };
    a.Reduce = [](const void *s, void *r, size_t n, int t, int, int root, ncclComm_t c, cudaStream_t) { return emu_push(EmuOp{0, s, r, n, t, root, c->rank}); };
    a.Send = [](const void *s, size_t n, int t, int peer, ncclComm_t c, cudaStream_t) { return emu_push(EmuOp{1, s, nullptr, n, t, peer, c->rank}); };
    a.Recv = [](void *r, size_t n, int t, int peer, ncclComm_t c, cudaStream_t) { return emu_push(EmuOp{2, nullptr, r, n, t, peer, c->rank}); };
    a.GetVersion = [](int *v) { *v = 0; return 0; };
    return true;
}
}  // namespace monte
