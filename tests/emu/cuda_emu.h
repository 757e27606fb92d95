// cuda_emu.h -- TEST INFRASTRUCTURE: a tiny single-process SIMT emulator.
//
// It lets the *unmodified* kernel sources under biogo_b200/csrc (*.cuh) be compiled with
// g++ and executed on the CPU, one fiber per CUDA thread, so the kernel logic (indexing,
// borders, shuffles, tie-break codes, traceback walk) can be checked against the oracle in
// the `-m "not gpu"` test tier of a container that has no GPU.  It is NOT a product path:
// libbiogoalign.so never links it, and the Python package never loads the emulated
// library (only tests/ do, by explicit path).
//
// Model: blocks run one after another; the threads of a block are ucontext fibers
// scheduled round-robin; warp-level primitives and __syncthreads are rendezvous points.
#pragma once

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <mutex>
#include <vector>

namespace emu {

struct dim3_ {
    unsigned x = 1, y = 1, z = 1;
};

struct Warp {
    uint64_t slot[32];
    int arrived = 0, gen = 0, live = 0;
};

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
    unsigned tid = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    int cur = 0;
    int barrier_arrived = 0, barrier_gen = 0, live = 0;
    ucontext_t sched;
    std::function<void()> body;
};

inline Block *&cur_block()
{
    static Block *b = nullptr;
    return b;
}
inline dim3_ &g_threadIdx()
{
    static dim3_ d;
    return d;
}
inline dim3_ &g_blockIdx()
{
    static dim3_ d;
    return d;
}
inline dim3_ &g_blockDim()
{
    static dim3_ d;
    return d;
}
inline dim3_ &g_gridDim()
{
    static dim3_ d;
    return d;
}

inline void yield()
{
    Block *b = cur_block();
    Fiber &f = b->fibers[b->cur];
    swapcontext(&f.ctx, &b->sched);
}

inline void fiber_entry()
{
    Block *b = cur_block();
    b->body();
    Fiber &f = b->fibers[b->cur];
    f.done = true;
    b->live--;
    b->warps[f.tid / 32].live--;
    // a finished lane must not block its warp's rendezvous
    Warp &w = b->warps[f.tid / 32];
    if (w.live > 0 && w.arrived == w.live) {
        w.arrived = 0;
        w.gen++;
    }
    if (b->live > 0 && b->barrier_arrived == b->live) {
        b->barrier_arrived = 0;
        b->barrier_gen++;
    }
    swapcontext(&f.ctx, &b->sched);
}

inline std::mutex &launch_mutex()
{
    static std::mutex m;
    return m;
}

inline void launch(unsigned grid, unsigned block, const std::function<void()> &body)
{
    // the emulator keeps its block state and every __shared__ array in statics: one launch at a time
    std::lock_guard<std::mutex> only_one(launch_mutex());
    const size_t STACK = 256 * 1024;
    Block b;
    b.body = body;
    b.fibers.resize(block);
    b.warps.resize((block + 31) / 32);
    for (unsigned t = 0; t < block; t++) b.fibers[t].stack = (char *)malloc(STACK);
    g_blockDim().x = block;
    g_gridDim().x = grid;
    cur_block() = &b;
    for (unsigned bid = 0; bid < grid; bid++) {
        g_blockIdx().x = bid;
        b.live = (int)block;
        b.barrier_arrived = 0;
        b.barrier_gen = 0;
        for (auto &w : b.warps) {
            w.arrived = 0;
            w.gen = 0;
            w.live = 0;
        }
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = b.fibers[t];
            f.done = false;
            f.tid = t;
            b.warps[t / 32].live++;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        while (b.live > 0) {
            for (unsigned t = 0; t < block; t++) {
                Fiber &f = b.fibers[t];
                if (f.done) continue;
                b.cur = (int)t;
                g_threadIdx().x = t;
                swapcontext(&b.sched, &f.ctx);
            }
        }
    }
    for (unsigned t = 0; t < block; t++) free(b.fibers[t].stack);
    cur_block() = nullptr;
}

inline Warp &my_warp()
{
    Block *b = cur_block();
    return b->warps[b->fibers[b->cur].tid / 32];
}
inline int my_lane() { return (int)(cur_block()->fibers[cur_block()->cur].tid & 31); }

inline void restore_tid()
{
    Block *b = cur_block();
    g_threadIdx().x = b->fibers[b->cur].tid;
}

inline void warp_barrier()
{
    Warp &w = my_warp();
    const int g = w.gen;
    if (++w.arrived >= w.live) {
        w.arrived = 0;
        w.gen++;
    } else {
        while (w.gen == g) yield();
    }
    restore_tid();
}

inline void block_barrier()
{
    Block *b = cur_block();
    const int g = b->barrier_gen;
    if (++b->barrier_arrived >= b->live) {
        b->barrier_arrived = 0;
        b->barrier_gen++;
    } else {
        while (b->barrier_gen == g) yield();
    }
    restore_tid();
}

template <typename T>
inline T exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Warp &w = my_warp();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[my_lane()] = raw;
    warp_barrier();
    T r = v;
    if (src_lane >= 0 && src_lane < 32) memcpy(&r, &w.slot[src_lane], sizeof(T));
    warp_barrier();
    return r;
}

}  // namespace emu

// ---- CUDA spellings -------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define threadIdx (emu::g_threadIdx())
#define blockIdx (emu::g_blockIdx())
#define blockDim (emu::g_blockDim())
#define gridDim (emu::g_gridDim())

struct uint2 {
    unsigned x, y;
};
struct uint4 {
    unsigned x, y, z, w;
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct int2 {
    int x, y;
};
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct alignas(16) int4 {
    int x, y, z, w;
};
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
struct ulonglong2 {
    unsigned long long x, y;
};
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

using std::max;
using std::min;

template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int d, int width = 32)
{
    const int l = emu::my_lane();
    const int base = l & ~(width - 1);
    return emu::exchange(v, (l - base) >= d ? l - d : -1);
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    const int base = emu::my_lane() & ~(width - 1);
    return emu::exchange(v, base + (src & (width - 1)));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
    (void)width;
    return emu::exchange(v, emu::my_lane() ^ m);
}
static inline unsigned __ballot_sync(unsigned, bool pred)
{
    emu::Warp &w = emu::my_warp();
    w.slot[emu::my_lane()] = pred ? 1 : 0;
    emu::warp_barrier();
    unsigned r = 0;
    for (int l = 0; l < 32; l++)
        if (w.slot[l]) r |= 1u << l;
    emu::warp_barrier();
    return r;
}
static inline int __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline void __syncwarp() { emu::warp_barrier(); }
static inline void __syncthreads() { emu::block_barrier(); }
static inline int atomicAdd(int *p, int v)
{
    int o = *p;
    *p = o + v;
    return o;
}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v)
{
    unsigned long long o = *p;
    *p = o + v;
    return o;
}
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }

// DPX / SIMD-in-a-word intrinsics used by the kernels (scalar restatements)
static inline int __vimax3_s32(int a, int b, int c) { return std::max(a, std::max(b, c)); }
static inline int __vimax3_s32_relu(int a, int b, int c) { return std::max(0, std::max(a, std::max(b, c))); }
static inline int __viaddmax_s32(int a, int b, int c) { return std::max(a + b, c); }
static inline int __viaddmax_s32_relu(int a, int b, int c) { return std::max(0, std::max(a + b, c)); }
static inline unsigned __vimin3_u32(unsigned a, unsigned b, unsigned c) { return std::min(a, std::min(b, c)); }

// ---- a minimal CUDA *runtime* so biogo_b200/csrc/ba_api.cu (the host scheduler) builds
// and runs unchanged on the CPU: device memory is host memory, streams are synchronous.
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp {
    int major = 10, minor = 0, multiProcessorCount = 2;
};
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    *p = cudaDeviceProp();
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned)
{
    *s = (void *)1;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int)
{
    *s = (void *)1;
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi)
{
    *lo = 0;
    *hi = -1;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e)
{
    *e = (void *)1;
    return cudaSuccess;
}
enum { cudaEventDefault = 0, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned)
{
    *e = (void *)1;
    return cudaSuccess;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n)
{
    *n = 2;  // the emulator pretends to be a two-device box (multi-device scheduler tests)
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }   // everything runs synchronously here
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t)
{
    *ms = 0.f;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = calloc(1, n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p)
{
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMallocHost(void **p, size_t n)
{
    *p = malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFreeHost(void *p)
{
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t)
{
    memcpy(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t)
{
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t)
{
    *f = *t = (size_t)1 << 30;
    return cudaSuccess;
}
template <typename F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t)
{
    *n = 1;
    return cudaSuccess;
}

// ---- packed 16-bit SIMD-in-a-word intrinsics (scalar restatements) and PRMT
static inline short emu_lo(unsigned v) { return (short)(v & 0xffffu); }
static inline short emu_hi(unsigned v) { return (short)(v >> 16); }
static inline unsigned emu_pk(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
static inline unsigned __vadd2(unsigned a, unsigned b) { return emu_pk(emu_lo(a) + emu_lo(b), emu_hi(a) + emu_hi(b)); }
static inline unsigned __vsub2(unsigned a, unsigned b) { return emu_pk(emu_lo(a) - emu_lo(b), emu_hi(a) - emu_hi(b)); }
static inline unsigned __vmaxs2(unsigned a, unsigned b)
{
    return emu_pk(std::max(emu_lo(a), emu_lo(b)), std::max(emu_hi(a), emu_hi(b)));
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline unsigned __vimax_s16x2_relu(unsigned a, unsigned b) { return __vmaxs2(__vmaxs2(a, b), 0u); }
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vadd2(a, b), c); }
static inline unsigned __viaddmax_s16x2_relu(unsigned a, unsigned b, unsigned c)
{
    return __vmaxs2(__vmaxs2(__vadd2(a, b), c), 0u);
}
static inline unsigned __vibmax_s16x2(unsigned a, unsigned b, bool *ph, bool *pl)
{
    *ph = emu_hi(a) >= emu_hi(b);
    *pl = emu_lo(a) >= emu_lo(b);
    return __vmaxs2(a, b);
}
// prmt.b32 (default mode): selector nibble bit 3 replicates the sign of the selected byte
static inline unsigned emu_prmt(unsigned a, unsigned b, unsigned s)
{
    const uint64_t src = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int k = 0; k < 4; k++) {
        const unsigned nib = (s >> (4 * k)) & 0xf;
        unsigned byte = (unsigned)(src >> (8 * (nib & 7))) & 0xff;
        if (nib & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * k);
    }
    return r;
}
