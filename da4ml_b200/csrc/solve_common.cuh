// solve_common.cuh -- shared pieces of the persistent solve kernel: argmax candidate type, per-CTA context,
// column-list references, the group barrier / all-gather exchange, op-record access, histogram append.
#pragma once
#include "cmvm_num.cuh"
#include "cmvm_types.cuh"

namespace da {

// ------------------------------------------------------------------------------------------------
// small device helpers

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
#ifdef DA_CPU_SIM
    simt::poll_yield(); // a poll: let the other simulated threads run
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}

// atomic add on a counter of the CTA's shared-memory context.  The context is reached through a generic pointer, for
// which atomicAdd compiles to a generic-address ATOM; the explicit state space gives the native shared-memory atomic.
#ifndef DA_SMEM_ATOM
#define DA_SMEM_ATOM 1
#endif
__device__ __forceinline__ int smem_add(int *p, int v) {
#if DA_SMEM_ATOM && !defined(DA_CPU_SIM)
    int old;
    asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    return old;
#else
    return atomicAdd(p, v);
#endif
}

// A store that races with other threads BY DESIGN: either every racer writes the same value (dirty flags) or the readers
// accept the old and the new value alike (rewrite stamps of c0 / c1 while other CTAs purge them explicitly).  A plain
// store on the GPU; an atomic one under the CPU simulation so that its race checker reports only unintended races.
template <class T> __device__ __forceinline__ void st_racy(T *p, T v) {
#ifdef DA_CPU_SIM
    __atomic_store_n(p, v, __ATOMIC_RELAXED);
#else
    *p = v;
#endif
}

// the matching load: a value other threads may be changing right now (a fresh read every time it is executed)
template <class T> __device__ __forceinline__ T ld_racy(const T *p) {
#ifdef DA_CPU_SIM
    return __atomic_load_n(p, __ATOMIC_RELAXED);
#else
    return *(const volatile T *)p;
#endif
}

// A team of warps of the CTA that works through a phase together: the whole CTA (bar 0 = __syncthreads) or a part of it
// with its own hardware barrier (bar.sync id, threads), so that two independent phases can run side by side.
struct Team {
    int tid, nt; // this thread's index in the team, threads of the team (a multiple of 32)
    int bar;     // 0: the whole CTA
};
__device__ __forceinline__ Team team_all() { return Team{(int)threadIdx.x, (int)blockDim.x, 0}; }
__device__ __forceinline__ void team_sync(const Team &t) {
    if (t.bar == 0)
        __syncthreads();
    else {
#ifdef DA_CPU_SIM
        da_sim_named_barrier(t.bar, t.nt);
#else
        asm volatile("bar.sync %0, %1;" ::"r"(t.bar), "r"(t.nt) : "memory");
#endif
    }
}

struct Best {
    uint32_t score, khi, klo;
};
__device__ __forceinline__ bool best_gt(const Best &a, const Best &b) {
    if (a.score != b.score)
        return a.score > b.score;
    if (a.khi != b.khi)
        return a.khi > b.khi;
    return a.klo > b.klo;
}
__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Best o;
        o.score = __shfl_xor_sync(0xffffffffu, b.score, off);
        o.khi = __shfl_xor_sync(0xffffffffu, b.khi, off);
        o.klo = __shfl_xor_sync(0xffffffffu, b.klo, off);
        if (best_gt(o, b))
            b = o;
    }
    return b;
}


#define DA_HOT_MAX 64
// Block-level context kept in shared memory
struct BlockCtx {
    Best warp_best[32];
    int warp_sum[32];
    int warp_st[32];
    Best chosen;       // pair selected for the current step (score==0 -> none)
    int seg_len;       // entries (live + dead) in the common log of this CTA's histogram segment, [0, cap0)
    // Hot regions (owner-partitioned kernel): the entries whose passive partner is one of the first hot_n inputs this CTA
    // owns get a region of their own, [cap0 + k * hot_cap, +hot_cap).  Such an input is rewritten over and over, and every
    // time ALL entries of its region die at once: the region is simply reset instead of re-reading the chunks they led.
    int cap0, hot_n, hot_cap;
    int hot_len[DA_HOT_MAX];  // entries of each region (may run past hot_cap while a step appends: readers clamp)
    int hot_snap[DA_HOT_MAX]; // the same at the start of the step (what the argmax caches have to cover)
    Best pend[DA_HOT_MAX + 1]; // maximum of the entries appended to each log in the current step (0: the common log)
    int n_new;         // entries appended in the current step
    int live_old;      // live entries counted by the last full rescan (accounting mode)
    int n_dirty;       // chunks to re-read in the current step
    int status;        // sticky error
    int list_max;      // longest column list seen by this CTA
    unsigned long long r_count;   // digit pairs enumerated by this CTA (all steps)
    unsigned long long rescanned; // histogram entries re-read by this CTA (all steps)
    unsigned bar_target;
    unsigned epoch;    // exchanges done by this group so far (stamps the all-gather slots)
    int scratch_i[4];
    long long phase[8];
    long long t_last;
    long long poll_iters;
    long long peak[8];
    long long nslow[8];
    int cmp_out;
    int r_step;        // digit pairs enumerated in the current step
    int rescan_step;   // histogram entries re-read in the current step
    unsigned long long xw0[304], xw1[304], xw2[304]; // payload words gathered from every CTA of the group
    int xprefix[308];
};

struct Ctx {
    LaunchCfg cfg;
    int rank;
    GroupWs ws;
    FEnt *seg; // this CTA's histogram segment (global)
    // shared memory
    BlockCtx *b;
    uint32_t *cb_score, *cb_khi, *cb_klo; // per-chunk cached maximum
    unsigned char *cb_dirty;
    int *dirty_list;
};

// Barrier across the G CTAs of a group: monotonic counter, release on arrive / acquire on poll, split in
// arrive / wait so independent work overlaps the wait.  Cross-CTA data is always read with ld.cg.
__device__ __forceinline__ void group_arrive(const Ctx &cx) {
    __syncthreads();
    if (cx.cfg.G > 1 && threadIdx.x == 0) {
#ifdef DA_CPU_SIM
        __atomic_fetch_add(cx.ws.barrier, 1u, __ATOMIC_RELEASE);
#else
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cx.ws.barrier) : "memory");
#endif
        cx.b->bar_target += (unsigned)cx.cfg.G;
    }
}
__device__ __forceinline__ void group_wait(const Ctx &cx) {
    if (cx.cfg.G > 1 && threadIdx.x == 0) {
        const unsigned target = cx.b->bar_target;
        while ((int)(ld_acquire_u32(cx.ws.barrier) - target) < 0) {
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void group_sync(const Ctx &cx) {
    group_arrive(cx);
    group_wait(cx);
}

// All-gather exchange: a group barrier that also carries three 64-bit payload words per CTA.
// publish (one thread, after a __syncthreads): store the payload in this CTA's slot, fence, arrive on the group
// counter.  collect: ONE thread per CTA polls the counter (all CTAs polling all slots would hammer a single L2
// slice), then the first G threads read the slots.  Slots are double-buffered by exchange parity: a slot is
// overwritten two exchanges later, which no CTA can reach before every CTA has finished reading it.
#define DA_PAY_MASK 0xffffffffffffULL
__device__ __forceinline__ void xchg_publish(const Ctx &cx, unsigned long long p0, unsigned long long p1, unsigned long long p2) {
    BlockCtx &b = *cx.b;
    b.epoch += 1u;
    if (cx.cfg.G > 1) {
        unsigned long long *s = cx.ws.xchg + ((size_t)(b.epoch & 1u) * cx.cfg.G + cx.rank) * 4;
        __stcg(s + 0, p0);
        __stcg(s + 1, p1);
        __stcg(s + 2, p2);
        // release: (with the preceding bar.sync) every earlier write of the CTA, including the slot
#ifdef DA_CPU_SIM
        __atomic_fetch_add(cx.ws.barrier, 1u, __ATOMIC_ACQ_REL);
#else
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(cx.ws.barrier) : "memory");
#endif
        b.bar_target += (unsigned)cx.cfg.G;
    }
    else {
        b.xw0[0] = p0;
        b.xw1[0] = p1;
        b.xw2[0] = p2;
    }
}
// block-wide; on return b.xw0/1/2[0..G) hold every CTA's payload
__device__ __forceinline__ void xchg_collect(const Ctx &cx) {
    BlockCtx &b = *cx.b;
    if (cx.cfg.G > 1) {
        if (threadIdx.x == 0) {
            const unsigned target = b.bar_target;
            const long long t0 = clock64();
            int iters = 0;
            while ((int)(ld_acquire_u32(cx.ws.barrier) - target) < 0) {
                ++iters;
            }
            b.phase[6] += clock64() - t0; // pure polling time
            b.poll_iters += iters;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cx.cfg.G; i += blockDim.x) {
            const unsigned long long *s = cx.ws.xchg + ((size_t)(b.epoch & 1u) * cx.cfg.G + i) * 4;
            b.xw0[i] = __ldcg(s + 0);
            b.xw1[i] = __ldcg(s + 1);
            b.xw2[i] = __ldcg(s + 2);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void load_op(const ProblemDesc &p, uint32_t id, QInt &q, float &lat) {
    if ((int)id < p.n_in) {
        q.min = p.qint[3 * id + 0];
        q.max = p.qint[3 * id + 1];
        q.step = p.qint[3 * id + 2];
        lat = p.lat[id];
    }
    else {
        float4 v = __ldcg(&p.op_q[id]);
        q.min = v.x;
        q.max = v.y;
        q.step = v.z;
        lat = v.w;
    }
}

#define DA_DEAD 0xffffffffu
#define DA_LAP(k)                                                              \
    if (threadIdx.x == 0) {                                                    \
        const long long _now = clock64();                                      \
        cx.b->phase[k] += _now - cx.b->t_last;                                 \
        cx.b->peak[k] = max(cx.b->peak[k], _now - cx.b->t_last);               \
        if (_now - cx.b->t_last > 20000)                                       \
            cx.b->nslow[k] += 1;                                               \
        cx.b->t_last = _now;                                                   \
    }

// Chunks of the segment in use, and the entries [lo, hi) of chunk c (empty when the chunk lies beyond its region's
// fill).  `seg_len` = fill of the common log; the hot regions' fills come from their step-start snapshot.
__device__ __forceinline__ int chunks_in_use(const Ctx &cx, int seg_len) {
    const BlockCtx &b = *cx.b;
    const int cl = cx.cfg.chunk_log;
    return b.hot_n ? (b.cap0 + b.hot_n * b.hot_cap) >> cl : (seg_len + (1 << cl) - 1) >> cl;
}
__device__ __forceinline__ void chunk_span(const Ctx &cx, int c, int seg_len, int &lo, int &hi) {
    const BlockCtx &b = *cx.b;
    lo = c << cx.cfg.chunk_log;
    const int end = lo + (1 << cx.cfg.chunk_log);
    if (b.hot_n == 0 || lo < b.cap0)
        hi = min(end, seg_len);
    else {
        const int k = (lo - b.cap0) / b.hot_cap;
        hi = k < b.hot_n ? min(end, b.cap0 + k * b.hot_cap + b.hot_snap[k]) : lo;
    }
    hi = max(hi, lo);
}

// Append one histogram entry (created at step `stamp`) to this CTA's segment and fold it into the
// thread's running best.
__device__ __forceinline__ void
emit_entry(const ProblemDesc &p, const Ctx &cx, uint32_t lo, uint32_t hi, int shift, int sub, uint32_t count, QInt q0, float l0, QInt q1, float l1, uint32_t stamp, uint32_t thresh, Best &best, bool mark_dirty = true, int hot = -1) {
    uint32_t score;
    if (!pair_score(p.method, count, q0, l0, q1, l1, score))
        return; // NaN score: can never be selected
    const uint64_t key = pack_key(lo, hi, shift, sub);
    int pos = -1;
    if (hot >= 0) { // the passive partner's own region; when that is full the entry goes to the common log like any other
        const int q = smem_add(&cx.b->hot_len[hot], 1);
        if (q < cx.b->hot_cap)
            pos = cx.b->cap0 + hot * cx.b->hot_cap + q;
    }
    if (pos < 0) {
        pos = smem_add(&cx.b->seg_len, 1);
        if (pos >= cx.b->cap0) {
            cx.b->status = ST_FSEG_OVERFLOW;
            return;
        }
    }
    if (cx.cfg.accounting)
        smem_add(&cx.b->n_new, 1); // only the exact live count of accounting mode needs it
    FEnt e;
    e.x = score;
    e.y = stamp;
    e.z = (uint32_t)key;
    e.w = (uint32_t)(key >> 32);
    cx.seg[pos] = e;
    if (mark_dirty) // (else the caller marks the chunks of everything it appended)
        st_racy(&cx.cb_dirty[pos >> cx.cfg.chunk_log], (unsigned char)1); // this chunk's cached maximum does not cover the new entry yet
    if (score >= thresh) {
        Best c{score, e.w, e.z};
        if (best_gt(c, best))
            best = c;
    }
}


} // namespace da
