// cuda_emul.h -- executes the CUDA sources of splashsurf_b200/csrc on the CPU.  TEST INFRASTRUCTURE ONLY.
//
// The build container has no GPU, so this header lets g++ compile ss_pipeline.cu (+ ss_kernels.cuh, ss_post.cuh) unchanged:
//
//     g++ -x c++ -DSS_HOST_EMUL -include tests/emul/cuda_emul.h splashsurf_b200/csrc/ss_pipeline.cu -> libsplashsurf_emul.so
//
// * every CUDA thread of a block runs as a fiber (a six-register stack switch on x86-64, ucontext elsewhere); __syncthreads / __ballot_sync / __shfl_*_sync / __all_sync /
//   __syncwarp suspend the fiber until the other live threads of the block / warp have arrived, exactly the semantics the
//   kernels rely on (a divergent or missing collective dead-locks the scheduler and aborts with a message);
// * blocks of a launch are spread over host threads; __shared__ is thread-local static storage of the worker;
// * the CUDA runtime calls the host code makes are served from host memory; cub's device-wide sorts / scans are std:: ones.
//
// The result is the same C ABI computed by the same statements, slowly.  tests/test_emulated_pipeline.py runs small parity
// cases through it; it is never loaded by the splashsurf_b200 package (which has no CPU path) and never measured.
#pragma once
#include <cuda_runtime.h>      // host-side declarations: vector types, cudaError_t, the runtime API prototypes
#include <ucontext.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <thread>
#include <type_traits>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#undef __shared__
#undef __constant__
#undef __restrict__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __restrict__

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

// ------------------------------------------------------------------ fibers + collectives ----
namespace emul {
constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 256 * 1024;
enum State { READY, WARP_WAIT, BLOCK_WAIT, DONE };
enum Op { OP_BALLOT, OP_ALL, OP_ANY, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_SYNCWARP, OP_BAR, OP_BAR_OR };

// Context switch: on x86-64 a six-register stack switch (no signal-mask system calls); ucontext elsewhere.
#if defined(__x86_64__) && !defined(SS_EMUL_UCONTEXT)
#define SS_EMUL_FAST_SWITCH 1
extern "C" void ss_emul_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl ss_emul_switch
.type ss_emul_switch,@function
ss_emul_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size ss_emul_switch,.-ss_emul_switch
)");
#endif

struct Fiber {
#ifdef SS_EMUL_FAST_SWITCH
    void *sp;
#else
    ucontext_t ctx;
#endif
    State state;
    int op, pred, arg;
    unsigned long long val, res;
};
struct Cta {
#ifdef SS_EMUL_FAST_SWITCH
    void *sched_sp = nullptr;
#else
    ucontext_t sched;
#endif
    Fiber *f = nullptr;
    char *stacks = nullptr;
    int n = 0, cur = 0;
    void (*entry)(void *) = nullptr;
    void *entry_arg = nullptr;
    Cta() {
        f = new Fiber[kMaxThreads];
        stacks = (char *)mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (stacks == (char *)MAP_FAILED) { perror("emul: mmap"); abort(); }
    }
    ~Cta() { delete[] f; munmap(stacks, kStack * kMaxThreads); }
};
static thread_local Cta *g_cta = nullptr;

static inline void to_scheduler(Cta *C, Fiber &F) {
#ifdef SS_EMUL_FAST_SWITCH
    ss_emul_switch(&F.sp, C->sched_sp);
#else
    swapcontext(&F.ctx, &C->sched);
#endif
}
static inline void to_fiber(Cta *C, Fiber &F) {
#ifdef SS_EMUL_FAST_SWITCH
    ss_emul_switch(&C->sched_sp, F.sp);
#else
    swapcontext(&C->sched, &F.ctx);
#endif
}
static void trampoline() {
    Cta *C = g_cta;
    C->entry(C->entry_arg);
    C->f[C->cur].state = DONE;
#ifdef SS_EMUL_FAST_SWITCH
    to_scheduler(C, C->f[C->cur]);                              // a finished fiber is never resumed
    abort();
#endif
}

[[noreturn]] static void deadlock(Cta *C) {
    fprintf(stderr, "emul: dead-lock in block (%u,%u,%u): thread states:", blockIdx.x, blockIdx.y, blockIdx.z);
    for (int t = 0; t < C->n; ++t) if (C->f[t].state != DONE) fprintf(stderr, " %d:%s/%d", t, C->f[t].state == WARP_WAIT ? "warp" : "block", C->f[t].op);
    fprintf(stderr, "\n");
    abort();
}

static void resolve_warp(Cta *C, int lo, int hi) {
    int first = lo;
    while (C->f[first].state != WARP_WAIT) ++first;              // the caller guarantees at least one waiting lane
    const int want = C->f[first].op;
    unsigned ballot = 0; bool all = true, any = false;
    for (int t = lo; t < hi; ++t) {
        Fiber &F = C->f[t];
        if (F.state != WARP_WAIT) continue;
        if (F.op != want) { fprintf(stderr, "emul: lanes of one warp wait in different collectives (%d vs %d)\n", F.op, want); abort(); }
        if (F.pred) ballot |= 1u << (t - lo);
        all = all && F.pred; any = any || F.pred;
    }
    for (int t = lo; t < hi; ++t) {
        Fiber &F = C->f[t];
        if (F.state != WARP_WAIT) continue;
        const int lane = t - lo;
        int src = lane;
        switch (want) {
            case OP_BALLOT: F.res = ballot; break;
            case OP_ALL: F.res = all; break;
            case OP_ANY: F.res = any; break;
            case OP_SHFL: src = F.arg & 31; break;
            case OP_SHFL_UP: src = lane - F.arg; break;
            case OP_SHFL_DOWN: src = lane + F.arg; break;
            case OP_SHFL_XOR: src = lane ^ F.arg; break;
            default: break;
        }
        if (want >= OP_SHFL && want <= OP_SHFL_XOR) {
            const bool ok = src >= 0 && src < 32 && lo + src < hi && C->f[lo + src].state == WARP_WAIT;
            F.res = ok ? C->f[lo + src].val : F.val;
        }
    }
    for (int t = lo; t < hi; ++t) if (C->f[t].state == WARP_WAIT) C->f[t].state = READY;
}

static void run_block(Cta *C, int nthreads, unsigned bdx, unsigned bdy) {
    C->n = nthreads;
    for (int t = 0; t < nthreads; ++t) {
        Fiber &F = C->f[t];
#ifdef SS_EMUL_FAST_SWITCH
        // initial frame: six callee-saved register slots, then the entry address that the first switch "returns" to;
        // after that `ret` the stack pointer is 8 below a 16-byte boundary, as at any function entry
        uintptr_t top = ((uintptr_t)(C->stacks + (size_t)(t + 1) * kStack)) & ~(uintptr_t)15;
        void **frame = (void **)(top - 16 - 6 * sizeof(void *));
        for (int q = 0; q < 6; ++q) frame[q] = nullptr;
        frame[6] = (void *)&trampoline;
        F.sp = (void *)frame;
#else
        getcontext(&F.ctx);
        F.ctx.uc_stack.ss_sp = C->stacks + (size_t)t * kStack;
        F.ctx.uc_stack.ss_size = kStack;
        F.ctx.uc_link = &C->sched;
        makecontext(&F.ctx, trampoline, 0);
#endif
        F.state = READY;
    }
    int live = nthreads;
    while (live) {
        bool progressed = false;
        for (int t = 0; t < nthreads; ++t) {
            if (C->f[t].state != READY) continue;
            C->cur = t;
            threadIdx.x = (unsigned)t % bdx; threadIdx.y = ((unsigned)t / bdx) % bdy; threadIdx.z = (unsigned)t / (bdx * bdy);
            to_fiber(C, C->f[t]);
            if (C->f[t].state == DONE) --live;
            progressed = true;
        }
        // warp collectives: every live lane of the warp has arrived
        int block_wait = 0;
        for (int lo = 0; lo < nthreads; lo += 32) {
            const int hi = std::min(lo + 32, nthreads);
            int alive = 0, waiting = 0;
            for (int t = lo; t < hi; ++t) { alive += C->f[t].state != DONE; waiting += C->f[t].state == WARP_WAIT; block_wait += C->f[t].state == BLOCK_WAIT; }
            if (alive && waiting == alive) { resolve_warp(C, lo, hi); progressed = true; }
        }
        // block barrier: every live thread has arrived
        if (live && block_wait == live) {
            int any = 0;
            for (int t = 0; t < nthreads; ++t) if (C->f[t].state == BLOCK_WAIT) any |= C->f[t].pred;
            for (int t = 0; t < nthreads; ++t) if (C->f[t].state == BLOCK_WAIT) { C->f[t].res = (unsigned long long)any; C->f[t].state = READY; }
            progressed = true;
        }
        if (live && !progressed) deadlock(C);
    }
}

static inline unsigned long long collective(State st, int op, unsigned long long val, int arg, int pred) {
    Cta *C = g_cta;
    Fiber &F = C->f[C->cur];
    F.op = op; F.val = val; F.arg = arg; F.pred = pred; F.state = st;
    to_scheduler(C, F);
    return F.res;
}

static Cta *cta_pool(int slot) {
    static std::vector<Cta *> pool;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)pool.size() <= slot) pool.resize(slot + 1, nullptr);
    if (!pool[slot]) pool[slot] = new Cta();
    return pool[slot];
}
static int g_workers = 0;
static inline int workers() {
    if (!g_workers) {
        const char *e = getenv("SS_EMUL_THREADS");
        g_workers = e ? std::max(1, atoi(e)) : (int)std::max(1u, std::thread::hardware_concurrency());
    }
    return g_workers;
}

template <typename Body>
static void run_grid(dim3 grid, dim3 block, Body &body) {
    const uint64_t nblocks = (uint64_t)grid.x * grid.y * grid.z;
    const int nthreads = (int)(block.x * block.y * block.z);
    if (!nblocks || !nthreads) return;
    if (nthreads > kMaxThreads) { fprintf(stderr, "emul: block of %d threads\n", nthreads); abort(); }
    auto worker = [&](uint64_t first, uint64_t step) {
        Cta *mine = cta_pool((int)first);                    // one fiber set per worker slot, reused across launches
        g_cta = mine;
        mine->entry = [](void *p) { (*static_cast<Body *>(p))(); };
        mine->entry_arg = &body;
        blockDim = block; gridDim = grid;
        for (uint64_t b = first; b < nblocks; b += step) {
            blockIdx.x = (unsigned)(b % grid.x); blockIdx.y = (unsigned)((b / grid.x) % grid.y); blockIdx.z = (unsigned)(b / ((uint64_t)grid.x * grid.y));
            run_block(mine, nthreads, block.x, block.y);
        }
    };
    const int W = (int)std::min<uint64_t>((uint64_t)workers(), nblocks >= 8 ? nblocks / 4 : 1);
    if (W <= 1) { worker(0, 1); return; }
    std::vector<std::thread> pool;
    for (int w = 0; w < W; ++w) pool.emplace_back(worker, (uint64_t)w, (uint64_t)W);
    for (auto &t : pool) t.join();
}

template <typename... P, typename... A>
static void launch(dim3 grid, dim3 block, void (*kernel)(P...), A... args) {
    auto body = [&]() { kernel(args...); };
    run_grid(grid, block, body);
}
}  // namespace emul

#define SS_LAUNCH(kern, grid, block, stream, ...) emul::launch(dim3(grid), dim3(block), kern, __VA_ARGS__)
#define SS_LAUNCH_DYN(kern, grid, block, smem, stream, ...) emul::launch(dim3(grid), dim3(block), kern, __VA_ARGS__)   // dynamic shared memory: a static worker-local array (ss_exact.cuh)

// ------------------------------------------------------------------ device intrinsics ----
static inline void __syncthreads() { emul::collective(emul::BLOCK_WAIT, emul::OP_BAR, 0, 0, 0); }
static inline int __syncthreads_or(int p) { return (int)emul::collective(emul::BLOCK_WAIT, emul::OP_BAR_OR, 0, 0, p != 0); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::collective(emul::WARP_WAIT, emul::OP_SYNCWARP, 0, 0, 0); }
static inline unsigned __ballot_sync(unsigned, int p) { return (unsigned)emul::collective(emul::WARP_WAIT, emul::OP_BALLOT, 0, 0, p != 0); }
static inline int __all_sync(unsigned, int p) { return (int)emul::collective(emul::WARP_WAIT, emul::OP_ALL, 0, 0, p != 0); }
static inline int __any_sync(unsigned, int p) { return (int)emul::collective(emul::WARP_WAIT, emul::OP_ANY, 0, 0, p != 0); }
template <typename T> static inline unsigned long long emul_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T emul_unbits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane, int = 32) { return emul_unbits<T>(emul::collective(emul::WARP_WAIT, emul::OP_SHFL, emul_bits(v), lane, 0)); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return emul_unbits<T>(emul::collective(emul::WARP_WAIT, emul::OP_SHFL_UP, emul_bits(v), (int)d, 0)); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return emul_unbits<T>(emul::collective(emul::WARP_WAIT, emul::OP_SHFL_DOWN, emul_bits(v), (int)d, 0)); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emul_unbits<T>(emul::collective(emul::WARP_WAIT, emul::OP_SHFL_XOR, emul_bits(v), m, 0)); }

// IEEE single operations; the TU is built with -ffp-contract=off -fno-fast-math, the volatile stores forbid excess precision
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
using std::max;
using std::min;
static inline int min(int a, unsigned b) { return a < (int)b ? a : (int)b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned long long min(unsigned long long a, unsigned b) { return a < b ? a : b; }

// global-memory atomics (blocks run on several host threads)
template <typename T> static inline T atomicAdd(T *p, T v) {
    if constexpr (std::is_floating_point<T>::value) {
        T old = *p, want;
        do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
        return old;
    } else {
        return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
    }
}
template <typename T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <typename T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <typename T> static inline T atomicMin(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <typename T> static inline T atomicMax(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ------------------------------------------------------------------ CUDA runtime served from host memory ----
struct EmulEvent { std::chrono::steady_clock::time_point t; };
extern "C" {
cudaError_t cudaGetDeviceCount(int *n) { *n = 16; return cudaSuccess; }    // every "device" is this host (multi-rank tests use LOCAL_RANK)
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
#ifdef SS_EMUL_GUARD
// Guarded allocator: the buffer ends (16-byte aligned) right in front of an inaccessible page, so that any read or write
// past the requested size faults immediately, and starts out filled with 0xCD (cudaMalloc returns uninitialised memory); a header in front keeps the mapping size.  DevBuf allocates without slack
// in this mode (ss_pipeline.cu), so "requested size" is what the host code really asked for.
cudaError_t cudaMalloc(void **p, size_t n) {
    const size_t page = 4096, need = ((n ? n : 1) + 15) & ~(size_t)15;
    const size_t body = (need + 64 + page - 1) / page * page;
    char *m = (char *)mmap(nullptr, body + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == (char *)MAP_FAILED) return cudaErrorMemoryAllocation;
    mprotect(m + body, page, PROT_NONE);
    char *user = m + body - need;
    ((size_t *)(user - 16))[0] = body + page; ((char **)(user - 16))[1] = m;
    memset(user, 0xCD, need);                                   // cudaMalloc does not zero: reads of uninitialised memory show up
    *p = user;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
    if (p) { char *user = (char *)p; const size_t len = ((size_t *)(user - 16))[0]; char *m = ((char **)(user - 16))[1]; munmap(m, len); }
    return cudaSuccess;
}
#else
cudaError_t cudaMalloc(void **p, size_t n) {
    *p = malloc(n ? n : 1);
    if (*p && getenv("SS_EMUL_POISON")) memset(*p, 0xCD, n ? n : 1);   // optional: make reads of uninitialised device memory visible
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
#endif
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemcpyToSymbol(const void *sym, const void *s, size_t n, size_t off, cudaMemcpyKind) { memcpy((char *)sym + off, s, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)1; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t) new EmulEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete (EmulEvent *)e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { ((EmulEvent *)e)->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(((EmulEvent *)b)->t - ((EmulEvent *)a)->t).count();
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = (size_t)6 << 30; *t = (size_t)8 << 30; return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    memset(a, 0, sizeof(*a)); a->type = cudaMemoryTypeUnregistered; a->hostPointer = (void *)p; return cudaSuccess;
}
}

// ------------------------------------------------------------------ cub's device-wide primitives ----
namespace cub {
struct DeviceRadixSort {
    template <typename K> static unsigned long long field(K k, int b, int e) {
        const unsigned long long v = (unsigned long long)k >> b;
        return (e - b) >= 64 ? v : (v & ((1ull << (e - b)) - 1ull));
    }
    template <typename K, typename V>
    static cudaError_t SortPairs(void *tmp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, int n, int b, int e, cudaStream_t = 0) {
        if (!tmp) { bytes = 1; return cudaSuccess; }
        std::vector<int> idx((size_t)n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return field(kin[x], b, e) < field(kin[y], b, e); });
        std::vector<K> ks((size_t)n); std::vector<V> vs((size_t)n);
        for (int i = 0; i < n; ++i) { ks[i] = kin[idx[i]]; vs[i] = vin[idx[i]]; }
        if (n) { memcpy(kout, ks.data(), sizeof(K) * n); memcpy(vout, vs.data(), sizeof(V) * n); }
        return cudaSuccess;
    }
    template <typename K>
    static cudaError_t SortKeys(void *tmp, size_t &bytes, const K *kin, K *kout, int n, int b, int e, cudaStream_t = 0) {
        if (!tmp) { bytes = 1; return cudaSuccess; }
        std::vector<K> ks(kin, kin + n);
        std::stable_sort(ks.begin(), ks.end(), [&](K x, K y) { return field(x, b, e) < field(y, b, e); });
        if (n) memcpy(kout, ks.data(), sizeof(K) * n);
        return cudaSuccess;
    }
};
struct DeviceScan {
    template <typename I, typename O>
    static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, const I *in, O *out, int n, cudaStream_t = 0) {
        if (!tmp) { bytes = 1; return cudaSuccess; }
        O acc = 0;
        for (int i = 0; i < n; ++i) { const O v = (O)in[i]; out[i] = acc; acc += v; }
        return cudaSuccess;
    }
    template <typename I, typename O>
    static cudaError_t InclusiveSum(void *tmp, size_t &bytes, const I *in, O *out, int n, cudaStream_t = 0) {
        if (!tmp) { bytes = 1; return cudaSuccess; }
        O acc = 0;
        for (int i = 0; i < n; ++i) { acc += (O)in[i]; out[i] = acc; }
        return cudaSuccess;
    }
};
}  // namespace cub
