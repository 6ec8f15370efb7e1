// simt.h -- a small host-side SIMT shim (TEST INFRASTRUCTURE ONLY): runs the solver's CUDA kernel sources on the CPU so
// that their logic can be checked against the oracle without a GPU.  Every CUDA thread is a fiber (ucontext) on one OS
// thread; __syncthreads / warp collectives are rendezvous points, spin-waits on global memory yield to the other
// fibers, atomics are plain read-modify-writes (nothing runs concurrently).  Deterministic, and a missing participant
// of a barrier or a full-mask warp collective shows up as a reported deadlock instead of undefined behaviour.
// It models functional behaviour only: no memory-model weakness, no timing.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#define DA_CPU_SIM 1
// Optional ThreadSanitizer build (-fsanitize=thread -DSIMT_TSAN): every fiber is announced to TSan as a logical thread
// and barriers / warp collectives as synchronisation, so a pair of conflicting accesses of the KERNEL code that no
// barrier orders is reported as a data race -- a race checker for shared and global memory alike.  The shim's own
// bookkeeping is excluded from the instrumentation.
#ifdef SIMT_TSAN
extern "C" {
void *__tsan_get_current_fiber(void);
void *__tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void *fiber);
void __tsan_switch_to_fiber(void *fiber, unsigned flags);
void __tsan_acquire(void *addr);
void __tsan_release(void *addr);
void __tsan_ignore_thread_begin(void);
void __tsan_ignore_thread_end(void);
}
#define SIMT_NOSAN __attribute__((no_sanitize("thread")))
namespace simt {
struct Quiet { // the shim's own bookkeeping (and the library code it calls) is not part of the program under test
    Quiet() { __tsan_ignore_thread_begin(); }
    ~Quiet() { __tsan_ignore_thread_end(); }
};
} // namespace simt
#define SIMT_QUIET simt::Quiet simt_quiet_scope
#else
#define SIMT_NOSAN
#define SIMT_QUIET
#endif
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct alignas(8) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace simt {

struct Fiber;
struct Barrier {
    int need = 0, count = 0;
    std::vector<Fiber *> waiters;
};
struct Warp {
    Barrier bar;
    uint64_t slot[2][32];
};
struct Cta {
    Barrier bar;
    Barrier named[16]; // bar.sync id, n
    std::vector<unsigned char> dyn;                              // dynamic shared memory
    std::map<int, std::unique_ptr<unsigned char[]>> statics;     // __shared__ variables, keyed by declaration
    std::vector<Warp> warps;
    int vote[2] = {0, 0}; // __syncthreads_or / _and / _count accumulators (double-buffered)
};
struct Fiber {
    ucontext_t ctx;
    std::unique_ptr<unsigned char[]> stack;
    bool done = false, blocked = false;
    uint3 tid{0, 0, 0}, bid{0, 0, 0};
    Cta *cta = nullptr;
    Warp *warp = nullptr;
    int lane = 0;
    unsigned collectives = 0; // warp collectives executed so far (selects the exchange buffer)
    unsigned votes = 0;       // block-wide votes executed so far
    void *tsan = nullptr;     // TSan's handle of this fiber
};
struct Grid {
    dim3 grid, block;
    std::vector<Cta> ctas;
    std::vector<Fiber> fibers;
    ucontext_t sched;
    Fiber *cur = nullptr;
    long long clock = 0;
    long long idle_polls = 0; // polls of global memory since the last barrier release / thread exit
    void *sched_tsan = nullptr;
    std::function<void()> body;
    std::string error;
};
// Order in which runnable fibers are resumed: 0 ascending thread ids, 1 descending, >= 2 a pseudo-random permutation that
// changes on every scheduler pass.  A kernel that is correct only under one order has a missing barrier.
inline int &schedule_mode() {
    static int m = 0;
    return m;
}
inline Grid *&current() {
    static Grid *g = nullptr;
    return g;
}
SIMT_NOSAN inline Fiber &self() { return *current()->cur; }
SIMT_NOSAN inline void to_scheduler() {
    Grid *g = current();
#ifdef SIMT_TSAN
    __tsan_switch_to_fiber(g->sched_tsan, 1); // 1 = no implicit synchronisation between the fibers
#endif
    swapcontext(&g->cur->ctx, &g->sched);
}
SIMT_NOSAN inline void yield() { to_scheduler(); }
// a spin-wait on global memory: give the other threads a turn; a wait nobody will ever satisfy is reported
SIMT_NOSAN inline void poll_yield() {
    SIMT_QUIET;
    Grid *g = current();
    if (++g->idle_polls > 20000000LL)
        throw std::runtime_error("simt livelock: a thread of block " + std::to_string(g->cur->bid.x) + " polls global memory while no other thread makes progress");
    yield();
}
// `orders_memory`: __syncthreads / __syncwarp order the participants' memory accesses; the rendezvous inside a shuffle or
// a vote does not (CUDA gives no memory ordering for them), so the race checker must not take it for one
SIMT_NOSAN inline void barrier_wait(Barrier &b, bool orders_memory = true) {
    SIMT_QUIET;
    Fiber *me = current()->cur;
#ifdef SIMT_TSAN
    if (orders_memory)
        __tsan_release(&b); // everything this thread did so far happens before whatever anybody does after the barrier
#endif
    if (++b.count == b.need) {
        b.count = 0;
        current()->idle_polls = 0;
        for (Fiber *f : b.waiters)
            f->blocked = false;
        b.waiters.clear();
    }
    else {
        me->blocked = true;
        b.waiters.push_back(me);
        yield();
    }
#ifdef SIMT_TSAN
    if (orders_memory)
        __tsan_acquire(&b);
#endif
}
SIMT_NOSAN inline void trampoline() {
    Grid *g = current();
#ifdef SIMT_TSAN
    __tsan_acquire(g); // the launch: everything the host did before it
#endif
    try {
        g->body();
    } catch (const std::exception &e) {
        SIMT_QUIET;
        g->error = e.what();
    }
#ifdef SIMT_TSAN
    __tsan_release(g); // kernel completion: visible to the host after launch() returns
#endif
    {
        SIMT_QUIET;
        g->cur->done = true;
        g->idle_polls = 0;
    }
    to_scheduler(); // (never resumed)
}

// Run `body` (a call of the kernel function with its arguments bound) on grid x block fibers.  blockDim.x must be a
// multiple of 32.  Throws on deadlock (a barrier / collective some participant never reaches) or on a kernel exception.
SIMT_NOSAN inline void launch(dim3 grid, dim3 block, size_t dyn_smem, std::function<void()> body, size_t stack_bytes = 512 * 1024) {
    SIMT_QUIET;
    if (block.x % 32 || block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1)
        throw std::runtime_error("simt::launch: 1-D launches with blockDim.x % 32 == 0 only");
    Grid g;
    g.grid = grid;
    g.block = block;
    g.body = std::move(body);
    g.ctas.resize(grid.x);
    g.fibers.resize((size_t)grid.x * block.x);
    const int nw = block.x / 32;
    for (unsigned b = 0; b < grid.x; ++b) {
        Cta &c = g.ctas[b];
        c.bar.need = (int)block.x;
        c.dyn.assign(dyn_smem + 16, 0);
        c.warps.resize(nw);
        for (auto &w : c.warps)
            w.bar.need = 32;
    }
    Grid *prev = current();
    current() = &g;
    for (unsigned b = 0; b < grid.x; ++b)
        for (unsigned t = 0; t < block.x; ++t) {
            Fiber &f = g.fibers[(size_t)b * block.x + t];
            f.tid = uint3{t, 0, 0};
            f.bid = uint3{b, 0, 0};
            f.cta = &g.ctas[b];
            f.warp = &g.ctas[b].warps[t / 32];
            f.lane = (int)(t % 32);
            f.stack.reset(new unsigned char[stack_bytes]);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.get();
            f.ctx.uc_stack.ss_size = stack_bytes;
            f.ctx.uc_link = &g.sched;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
#ifdef SIMT_TSAN
            f.tsan = __tsan_create_fiber(0);
#endif
        }
#ifdef SIMT_TSAN
    g.sched_tsan = __tsan_get_current_fiber();
    __tsan_release(&g);
#endif
    size_t remaining = g.fibers.size();
    std::vector<uint32_t> order(g.fibers.size());
    for (size_t i = 0; i < order.size(); ++i)
        order[i] = (uint32_t)(schedule_mode() == 1 ? order.size() - 1 - i : i);
    uint64_t rng = 0x9e3779b97f4a7c15ull * (uint64_t)(schedule_mode() + 1);
    while (remaining) {
        bool progressed = false;
        if (schedule_mode() >= 2)
            for (size_t i = order.size(); i > 1; --i) { // Fisher-Yates with xorshift
                rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17;
                std::swap(order[i - 1], order[rng % i]);
            }
        for (uint32_t fi : order) {
            Fiber &f = g.fibers[fi];
            if (f.done || f.blocked)
                continue;
            g.cur = &f;
#ifdef SIMT_TSAN
            __tsan_switch_to_fiber(f.tsan, 1);
#endif
            swapcontext(&g.sched, &f.ctx);
            progressed = true;
            if (f.done)
                --remaining;
            if (!g.error.empty())
                break;
        }
        if (!g.error.empty() || !progressed)
            break;
    }
    current() = prev;
#ifdef SIMT_TSAN
    __tsan_acquire(&g);
    for (Fiber &f : g.fibers)
        __tsan_destroy_fiber(f.tsan);
#endif
    if (!g.error.empty())
        throw std::runtime_error("simt kernel exception: " + g.error);
    if (remaining) {
        size_t blocked = 0;
        for (Fiber &f : g.fibers)
            blocked += !f.done && f.blocked;
        throw std::runtime_error("simt deadlock: " + std::to_string(remaining) + " threads unfinished, " + std::to_string(blocked) + " blocked at a barrier / warp collective");
    }
}

template <class T> SIMT_NOSAN T *shared_var(int key) {
    SIMT_QUIET;
    Cta &c = *self().cta;
    auto it = c.statics.find(key);
    if (it == c.statics.end()) {
        std::unique_ptr<unsigned char[]> p(new unsigned char[sizeof(T) + alignof(T)]());
        it = c.statics.emplace(key, std::move(p)).first;
    }
    uintptr_t a = (uintptr_t)it->second.get();
    a = (a + alignof(T) - 1) & ~(uintptr_t)(alignof(T) - 1);
    return (T *)a;
}
SIMT_NOSAN inline unsigned char *dyn_shared() {
    uintptr_t a = (uintptr_t)self().cta->dyn.data();
    return (unsigned char *)((a + 15) & ~(uintptr_t)15);
}

// warp-wide exchange of one value per lane: double-buffered, one rendezvous per collective
template <class T> SIMT_NOSAN T exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "exchange: values up to 64 bits");
    SIMT_QUIET;
    Fiber &me = self();
    Warp &w = *me.warp;
    const unsigned p = me.collectives++ & 1u;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[p][me.lane] = raw;
    barrier_wait(w.bar, false);
    T r;
    memcpy(&r, &w.slot[p][src_lane & 31], sizeof(T));
    return r;
}
SIMT_NOSAN inline unsigned ballot(bool pred) {
    SIMT_QUIET;
    Fiber &me = self();
    Warp &w = *me.warp;
    const unsigned p = me.collectives++ & 1u;
    w.slot[p][me.lane] = pred ? 1u : 0u;
    barrier_wait(w.bar, false);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l)
        m |= (unsigned)(w.slot[p][l] & 1u) << l;
    return m;
}

} // namespace simt

// ---- CUDA surface used by the kernel sources ----------------------------------------------------------------
#define threadIdx (simt::self().tid)
#define blockIdx (simt::self().bid)
#define blockDim (simt::current()->block)
#define gridDim (simt::current()->grid)
#define DA_SHARED_VAR(type, name) type &name = *simt::shared_var<type>(__COUNTER__)
#define DA_DYN_SHARED(name) unsigned char *name = simt::dyn_shared()

inline void __syncthreads() { simt::barrier_wait(simt::self().cta->bar); }
// bar.sync id, n: barrier `id` (1..15) of the CTA for the n threads that use it
SIMT_NOSAN inline void da_sim_named_barrier(int id, int n) {
    simt::Barrier &b = simt::self().cta->named[id & 15];
    {
        SIMT_QUIET;
        if (b.count == 0)
            b.need = n;
        else if (b.need != n)
            throw std::runtime_error("simt: named barrier used with two different thread counts");
    }
    simt::barrier_wait(b);
}
SIMT_NOSAN inline int __syncthreads_count(int pred) {
    SIMT_QUIET;
    simt::Fiber &me = simt::self();
    simt::Cta &c = *me.cta;
    const unsigned p = me.votes++ & 1u;
    c.vote[p] += pred ? 1 : 0;
    simt::barrier_wait(c.bar);
    const int r = c.vote[p];
    simt::barrier_wait(c.bar);
    c.vote[p] = 0; // (every thread, before it can reach the vote after next, which reuses this buffer)
    return r;
}
inline int __syncthreads_or(int pred) { return __syncthreads_count(pred) != 0; }
inline int __syncthreads_and(int pred) { return __syncthreads_count(pred) == (int)simt::current()->block.x; }
SIMT_NOSAN inline void __syncwarp(unsigned = 0xffffffffu) {
    ++simt::self().collectives; // keeps the buffer parity of the lanes aligned with exchange() / ballot()
    simt::barrier_wait(simt::self().warp->bar);
}
inline void sim_require_full(unsigned mask) {
    if (mask != 0xffffffffu)
        throw std::runtime_error("simt: partial-mask warp collectives are not modelled");
}
template <class T> T __shfl_sync(unsigned m, T v, int src) { sim_require_full(m); return simt::exchange(v, src); }
template <class T> T __shfl_xor_sync(unsigned m, T v, int off) { sim_require_full(m); return simt::exchange(v, simt::self().lane ^ off); }
template <class T> T __shfl_up_sync(unsigned m, T v, int d) {
    sim_require_full(m);
    const int lane = simt::self().lane;
    return simt::exchange(v, lane >= d ? lane - d : lane);
}
template <class T> T __shfl_down_sync(unsigned m, T v, int d) {
    sim_require_full(m);
    const int lane = simt::self().lane;
    return simt::exchange(v, lane + d < 32 ? lane + d : lane);
}
inline unsigned __ballot_sync(unsigned m, bool p) { sim_require_full(m); return simt::ballot(p); }
inline bool __any_sync(unsigned m, bool p) { sim_require_full(m); return simt::ballot(p) != 0u; }
inline bool __all_sync(unsigned m, bool p) { sim_require_full(m); return simt::ballot(p) == 0xffffffffu; }

// device atomics: real (relaxed) atomic builtins -- nothing runs concurrently, but a race checker must see them as atomics
template <class T, class U> T atomicAdd(T *p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_RELAXED); }
template <class T, class U> T atomicSub(T *p, U v) { return __atomic_fetch_sub(p, (T)v, __ATOMIC_RELAXED); }
template <class T, class U> T atomicOr(T *p, U v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_RELAXED); }
template <class T, class U> T atomicAnd(T *p, U v) { return __atomic_fetch_and(p, (T)v, __ATOMIC_RELAXED); }
template <class T, class U> T atomicExch(T *p, U v) { return __atomic_exchange_n(p, (T)v, __ATOMIC_RELAXED); }
template <class T, class U> T atomicMax(T *p, U v) {
    T o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while ((T)v > o && !__atomic_compare_exchange_n(p, &o, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return o;
}
template <class T, class U> T atomicMin(T *p, U v) {
    T o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while ((T)v < o && !__atomic_compare_exchange_n(p, &o, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return o;
}
template <class T, class U, class V> T atomicCAS(T *p, U c, V v) {
    T e = (T)c;
    __atomic_compare_exchange_n(p, &e, (T)v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return e;
}
// ld.cg is how the kernels read data other CTAs may be writing: a relaxed atomic load where the type allows it
template <class T> T __ldcg(const T *p) {
    if constexpr (sizeof(T) == 4 || sizeof(T) == 8) {
        typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type raw = __atomic_load_n((const typename std::conditional<sizeof(T) == 4, uint32_t, uint64_t>::type *)p, __ATOMIC_RELAXED);
        T v;
        memcpy(&v, &raw, sizeof(T));
        return v;
    }
    else
        return *p;
}
template <class T> T __ldg(const T *p) { return *p; }
template <class T, class U> void __stcg(T *p, U v) { *p = (T)v; }
inline void __threadfence() {}

// CUDA's global min / max overloads
template <class A, class B> inline typename std::common_type<A, B>::type min(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)b < (T)a ? (T)b : (T)a;
}
template <class A, class B> inline typename std::common_type<A, B>::type max(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)a < (T)b ? (T)b : (T)a;
}

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline unsigned __fns(unsigned mask, unsigned base, int offset) { // offset-th set bit at or above `base` (offset >= 1)
    if (offset < 1)
        throw std::runtime_error("simt: __fns is modelled for positive offsets only");
    for (unsigned b = base; b < 32; ++b)
        if ((mask >> b) & 1u)
            if (--offset == 0)
                return b;
    return 0xffffffffu;
}
inline int __float2int_rd(float x) { return (int)floorf(x); }
inline int __float2int_rn(float x) { return (int)nearbyintf(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
SIMT_NOSAN inline long long clock64() { return ++simt::current()->clock; }
