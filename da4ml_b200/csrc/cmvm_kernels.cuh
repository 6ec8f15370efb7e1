// cmvm_kernels.cuh -- sm_100a kernels of the CMVM greedy common-subexpression solver.
//
//   cmvm_prep_kernel    centre + CSD-decompose the constant matrix into packed sign planes
//                       (bit_decompose.hh:21-34, bit_decompose.cc:22-62, state_opr.cc:92-97)
//   cmvm_solve_kernel   persistent kernel; a group of G CTAs owns one problem:
//                       build column lists + initial pair histogram (state_opr.cc:100-144),
//                       greedy loop = select / substitute / recount (cmvm_core.cc:36-70,
//                       indexers.cc, state_opr.cc:227-345), adder-tree finisher (cmvm_core.cc:89-225)
//
// Formulation (differs from the reference by design, results are identical):
//   * a row of an expression in one output column is two 32-bit sign planes, so a pair count is a
//     handful of AND/shift/popc and substitution is a mask operation;
//   * the histogram is an unordered, append-only log of (score, stamp, packed key) split into one
//     segment per CTA.  The reference's "erase every entry touching id0/id1" (state_opr.cc:291-294)
//     is lazy: each expression carries the step at which it was last rewritten, an entry is live iff
//     its creation stamp is not older than either operand's rewrite step.  The argmax keeps a
//     per-chunk cached maximum in shared memory and re-reads a chunk only when its cached winner
//     died or entries were appended to it; order independence comes from reducing on the composite
//     (score, key), whose order is exactly the reference's "last maximum in sorted order";
//   * the live rows of the output columns a CTA owns are staged in shared memory (global memory
//     when they do not fit);
//   * recounting after a substitution enumerates digit pairs only in the columns that hold the
//     modified rows, accumulating into a zero-initialised counter slab with L2 atomics; the first
//     toucher of a counter records it, so harvesting costs O(distinct pairs) and leaves the slab zero.
#pragma once
#include "cmvm_num.cuh"
#include "cmvm_types.cuh"

namespace da {

// ------------------------------------------------------------------------------------------------
// small device helpers

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
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


// Block-level context kept in shared memory
struct BlockCtx {
    Best warp_best[32];
    int warp_sum[32];
    int warp_st[32];
    Best chosen;       // pair selected for the current step (score==0 -> none)
    int seg_len;       // entries (live + dead) in this CTA's histogram segment
    int n_new;         // entries appended in the current step
    int live_old;      // live entries counted by the last full rescan (accounting mode)
    int touch_n;       // counters first-touched by this CTA in the current step
    int n_act;         // owned columns touched by the current substitution
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

// one owned column touched by the current substitution (filled by the column's warp, read by the whole CTA)
struct ActCol {
    int o, slot;             // global column index, local slot
    int pos0, pos1, posn;    // list positions of the rows of c0, c1 and the new expression (-1: none)
    uint32_t P0, N0, P1, N1, Pn, Nn; // their sign planes after the substitution
};

struct ColRef {
    uint32_t *e, *P, *N; // structure-of-arrays list of one column: expression id and sign planes
    int *len;
    int cap;
};

struct Ctx {
    LaunchCfg cfg;
    int rank;
    GroupWs ws;
    FEnt *seg;          // this CTA's histogram segment (global)
    uint32_t *touch_g;  // overflow of the touched-counter list (global)
    // shared memory
    BlockCtx *b;
    uint32_t *cb_score, *cb_khi, *cb_klo; // per-chunk cached maximum
    unsigned char *cb_dirty;
    int *dirty_list;
    int *col_len_s;
    ActCol *act;
    uint32_t *lists_s;
};

__device__ __forceinline__ ColRef col_ref(const Ctx &cx, const ProblemDesc &p, int slot, int o) {
    ColRef r;
    if (cx.cfg.lcap > 0) {
        uint32_t *base = cx.lists_s + (size_t)slot * 3 * cx.cfg.lcap;
        r.e = base;
        r.P = base + cx.cfg.lcap;
        r.N = base + 2 * cx.cfg.lcap;
        r.len = &cx.col_len_s[slot];
        r.cap = cx.cfg.lcap;
    }
    else {
        uint32_t *base = cx.ws.col_u32 + (size_t)o * 3 * p.col_cap;
        r.e = base;
        r.P = base + p.col_cap;
        r.N = base + 2 * p.col_cap;
        r.len = &cx.ws.col_len[o];
        r.cap = p.col_cap;
    }
    return r;
}

// Barrier across the G CTAs of a group: monotonic counter, release on arrive / acquire on poll, split in
// arrive / wait so independent work overlaps the wait.  Cross-CTA data is always read with ld.cg.
__device__ __forceinline__ void group_arrive(const Ctx &cx) {
    __syncthreads();
    if (cx.cfg.G > 1 && threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cx.ws.barrier) : "memory");
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
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(cx.ws.barrier) : "memory");
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

// Append one histogram entry (created at step `stamp`) to this CTA's segment and fold it into the
// thread's running best.
__device__ __forceinline__ void
emit_entry(const ProblemDesc &p, const Ctx &cx, uint32_t lo, uint32_t hi, int shift, int sub, uint32_t count, QInt q0, float l0, QInt q1, float l1, uint32_t stamp, uint32_t thresh, Best &best) {
    uint32_t score;
    if (!pair_score(p.method, count, q0, l0, q1, l1, score))
        return; // NaN score: can never be selected
    const uint64_t key = pack_key(lo, hi, shift, sub);
    const int pos = atomicAdd(&cx.b->seg_len, 1);
    if (pos >= cx.ws.fseg_cap) {
        cx.b->status = ST_FSEG_OVERFLOW;
        return;
    }
    atomicAdd(&cx.b->n_new, 1);
    FEnt e;
    e.x = score;
    e.y = stamp;
    e.z = (uint32_t)key;
    e.w = (uint32_t)(key >> 32);
    cx.seg[pos] = e;
    cx.cb_dirty[pos >> cx.cfg.chunk_log] = 1; // this chunk's cached maximum does not cover the new entry yet
    if (score >= thresh) {
        Best c{score, e.w, e.z};
        if (best_gt(c, best))
            best = c;
    }
}

// ------------------------------------------------------------------------------------------------
// prep: one CTA per problem

__global__ void __launch_bounds__(256) cmvm_prep_kernel(ProblemDesc *probs) {
    ProblemDesc &p = probs[blockIdx.x];
    const int n_in = p.n_in, n_out = p.n_out;
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ int s_max, s_d0, s_colcap, s_dcolmax, s_rowsmax;
    if (tid == 0) {
        s_max = 0;
        s_d0 = 0;
        s_colcap = 0;
        s_dcolmax = 0;
        s_rowsmax = 0;
    }
    // column shifts (bit_decompose.hh:29): shift1[j] = min_i lsb(k[i,j])
    for (int j = tid; j < n_out; j += nt) {
        int m = 127;
        for (int i = 0; i < n_in; ++i)
            m = min(m, (int)get_lsb_loc(p.kernel[(size_t)i * n_out + j]));
        p.shift1[j] = (int8_t)m;
    }
    __syncthreads();
    // row shifts on the column-scaled matrix (bit_decompose.hh:31)
    for (int i = tid; i < n_in; i += nt) {
        int m = 127;
        for (int j = 0; j < n_out; ++j) {
            float v = (float)((double)p.kernel[(size_t)i * n_out + j] * exp2(-(double)p.shift1[j]));
            m = min(m, (int)get_lsb_loc(v));
        }
        p.shift0[i] = (int8_t)m;
    }
    __syncthreads();
    // global max |centred| -> CSD width (bit_decompose.cc:23-27)
    int lmax = 0;
    for (int idx = tid; idx < n_in * n_out; idx += nt) {
        int i = idx / n_out, j = idx - i * n_out;
        float v = (float)((double)p.kernel[idx] * exp2(-(double)p.shift1[j]));
        v = (float)((double)v * exp2(-(double)p.shift0[i]));
        int x = (int)v;
        lmax = max(lmax, abs(x));
    }
    atomicMax(&s_max, lmax);
    __syncthreads();
    int N = ceil_log2_pos((double)fmaxf((float)s_max, 1.0f) * 1.5);
    N = max(N, 1);
    // digits (bit_decompose.cc:29-38) -> sign planes; zero rows of zero-range inputs (state_opr.cc:92-97)
    for (int idx = tid; idx < n_in * n_out; idx += nt) {
        int i = idx / n_out, j = idx - i * n_out;
        float v = (float)((double)p.kernel[idx] * exp2(-(double)p.shift1[j]));
        v = (float)((double)v * exp2(-(double)p.shift0[i]));
        int x = (int)v;
        uint32_t P = 0, Nn = 0;
        for (int n = N - 1; n >= 0; --n) {
            int p2 = (int)(1u << n);
            int thres = (int)(((long long)p2 * 2) / 3);
            int d = (x > thres) - (x < -thres);
            if (d > 0)
                P |= 1u << n;
            if (d < 0)
                Nn |= 1u << n;
            x -= p2 * d;
        }
        if (p.qint[3 * i] == 0.0f && p.qint[3 * i + 1] == 0.0f) {
            P = 0;
            Nn = 0;
        }
        p.masks0[idx] = make_uint2(P, Nn);
    }
    __syncthreads();
    for (int j = tid; j < n_out; j += nt) {
        int d = 0, rows = 0;
        for (int i = 0; i < n_in; ++i) {
            uint2 m = p.masks0[(size_t)i * n_out + j];
            int c = __popc(m.x) + __popc(m.y);
            d += c;
            rows += (c != 0);
        }
        p.col_digits[j] = d;
        atomicAdd(&s_d0, d);
        atomicMax(&s_colcap, d + n_in);
        atomicMax(&s_dcolmax, d);
        atomicMax(&s_rowsmax, rows);
    }
    __syncthreads();
    if (tid == 0) {
        p.prep_meta[PM_NBITS] = N;
        p.prep_meta[PM_D0] = s_d0;
        p.prep_meta[PM_COLCAP] = s_colcap;
        p.prep_meta[PM_DCOL_MAX] = s_dcolmax;
        p.prep_meta[PM_ROWS_MAX] = s_rowsmax;
    }
}


// ------------------------------------------------------------------------------------------------
// solve: pieces

__device__ __forceinline__ uint32_t slab_index(const ProblemDesc &p, int slot, uint32_t x, int shift, int sub) {
    return (((uint32_t)slot * (uint32_t)p.e_cap + x) << p.log_s) + (uint32_t)(((shift + p.nbits - 1) << 1) | sub);
}

// first toucher of a counter records it: harvesting is O(distinct pairs) and leaves the slab zero
__device__ __forceinline__ void touch_push(const Ctx &cx, uint32_t idx) {
    const int t = atomicAdd(&cx.b->touch_n, 1);
    if (t < cx.ws.touch_cap)
        cx.touch_g[t] = idx;
    else
        cx.b->status = ST_TOUCH_OVERFLOW;
}
__device__ __forceinline__ void bump_now(const ProblemDesc &p, const Ctx &cx, int slot, uint32_t x, int shift, int sub) {
    const uint32_t idx = slab_index(p, slot, x, shift, sub);
    if (atomicAdd(&cx.ws.slab[idx], 1u) == 0u)
        touch_push(cx, idx);
}

// One source of digit pairs: every digit of row `lo` against every digit of row `hi` (state_opr.cc:331-336),
// enumerated by pair index so that a warp can walk all its lanes' pairs in lock step.
struct PairSrc {
    unsigned long long qlo, qhi; // digit positions of the two rows, 5 bits each, ascending (rows with <= 12 digits)
    uint32_t Plo, Nlo, Phi, Nhi;
    uint32_t base; // counter index of (slot, partner, shift = -(nbits-1), sub = 0)
    int dhi;       // digits in the hi row
    int n;         // number of pairs = digits(lo) * digits(hi)
    bool packed;   // qlo/qhi valid (else positions are found with __fns)
};
__device__ __forceinline__ unsigned long long pack_positions(uint32_t m) {
    unsigned long long q = 0ULL;
    int i = 0;
    for (; m; m &= m - 1, i += 5)
        q |= (unsigned long long)(__ffs(m) - 1) << i;
    return q;
}
__device__ __forceinline__ PairSrc make_src(const ProblemDesc &p, bool on, int slot, uint32_t x, uint32_t Plo, uint32_t Nlo, uint32_t Phi, uint32_t Nhi) {
    PairSrc s;
    s.Plo = Plo, s.Nlo = Nlo, s.Phi = Phi, s.Nhi = Nhi;
    s.base = ((uint32_t)slot * (uint32_t)p.e_cap + x) << p.log_s;
    const int dlo = __popc(Plo | Nlo);
    s.dhi = __popc(Phi | Nhi);
    s.n = on ? dlo * s.dhi : 0;
    s.packed = dlo <= 12 && s.dhi <= 12;
    s.qlo = s.qhi = 0ULL;
    if (s.n && s.packed) {
        s.qlo = pack_positions(Plo | Nlo);
        s.qhi = pack_positions(Phi | Nhi);
    }
    return s;
}
// counter index of the j-th pair of a source
__device__ __forceinline__ uint32_t pair_index(const ProblemDesc &p, const PairSrc &s, int j) {
    const int ja = __float2int_rd(__fdividef((float)j + 0.5f, (float)s.dhi)); // exact for these small integers
    const int jb = j - ja * s.dhi;
    int pl, ph;
    if (s.packed) {
        pl = (int)((s.qlo >> (5 * ja)) & 31ULL);
        ph = (int)((s.qhi >> (5 * jb)) & 31ULL);
    }
    else {
        pl = (int)__fns(s.Plo | s.Nlo, 0, ja + 1);
        ph = (int)__fns(s.Phi | s.Nhi, 0, jb + 1);
    }
    const int sub = (int)(((s.Nlo >> pl) ^ (s.Nhi >> ph)) & 1u);
    return s.base + (uint32_t)(((ph - pl + p.nbits - 1) << 1) | sub);
}
// digit pairs inside one row, state_opr.cc:323-330: v0 = higher digit, v1 = lower -> negative shift (rare: only
// the rewritten rows themselves)
__device__ __forceinline__ int pairs_self(const ProblemDesc &p, const Ctx &cx, int slot, uint32_t x, uint32_t P, uint32_t N) {
    int n = 0;
    for (uint32_t ma = P | N; ma; ma &= ma - 1) {
        const int pa = __ffs(ma) - 1;
        const int sa = (N >> pa) & 1;
        for (uint32_t mb = (P | N) & ((1u << pa) - 1u); mb; mb &= mb - 1) {
            const int pb = __ffs(mb) - 1;
            const int sb = (N >> pb) & 1;
            bump_now(p, cx, slot, x, pb - pa, sa ^ sb);
            ++n;
        }
    }
    return n;
}

// Substitution of the chosen pair inside one owned column, executed by one warp
// (update_expr, state_opr.cc:227-283).  Records the column in the CTA's active list when it holds
// one of the rewritten rows; the recount is done afterwards by the whole CTA.
__device__ void column_substitute(const ProblemDesc &p, const Ctx &cx, int slot, int o, uint32_t c0, uint32_t c1, int shift, int sub, uint32_t newid) {
    const int lane = threadIdx.x & 31;
    const ColRef L = col_ref(cx, p, slot, o);
    const int len = *L.len;

    int pos0 = -1, pos1 = -1;
    uint32_t P0 = 0, N0 = 0, P1 = 0, N1 = 0;
    for (int k = lane; k < len; k += 32) {
        const uint32_t e = L.e[k], P = L.P[k], N = L.N[k];
        if ((P | N) == 0)
            continue; // dead slot (keeps a stale id until it is recycled)
        if (e == c0) {
            pos0 = k;
            P0 = P;
            N0 = N;
        }
        if (e == c1) {
            pos1 = k;
            P1 = P;
            N1 = N;
        }
    }
    {
        const unsigned m0 = __ballot_sync(0xffffffffu, pos0 >= 0);
        const unsigned m1 = __ballot_sync(0xffffffffu, pos1 >= 0);
        const int s0 = m0 ? __ffs(m0) - 1 : 0, s1 = m1 ? __ffs(m1) - 1 : 0;
        pos0 = __shfl_sync(0xffffffffu, pos0, s0);
        P0 = __shfl_sync(0xffffffffu, P0, s0);
        N0 = __shfl_sync(0xffffffffu, N0, s0);
        pos1 = __shfl_sync(0xffffffffu, pos1, s1);
        P1 = __shfl_sync(0xffffffffu, P1, s1);
        N1 = __shfl_sync(0xffffffffu, N1, s1);
        if (!m0) {
            pos0 = -1;
            P0 = N0 = 0;
        }
        if (!m1) {
            pos1 = -1;
            P1 = N1 = 0;
        }
    }
    if (((P0 | N0) | (P1 | N1)) == 0)
        return; // neither operand lives in this column: nothing changes here

    uint32_t Pn = 0, Nn = 0;
    if (c0 != c1) {
        const bool flip = shift < 0;
        const int rel = flip ? -shift : shift;
        const uint32_t AP = flip ? P1 : P0, AN = flip ? N1 : N0; // expr0 after the reference's swap
        const uint32_t BP = flip ? P0 : P1, BN = flip ? N0 : N1;
        const uint32_t M = sub ? ((AP & (BN >> rel)) | (AN & (BP >> rel))) : ((AP & (BP >> rel)) | (AN & (BN >> rel)));
        const uint32_t MB = M << rel;
        const uint32_t AP2 = AP & ~M, AN2 = AN & ~M, BP2 = BP & ~MB, BN2 = BN & ~MB;
        if (!flip) { // the new digit takes position and sign of id0's digit
            Pn = AP & M;
            Nn = AN & M;
            P0 = AP2, N0 = AN2, P1 = BP2, N1 = BN2;
        }
        else {
            Pn = BP & MB;
            Nn = BN & MB;
            P1 = AP2, N1 = AN2, P0 = BP2, N0 = BN2;
        }
    }
    else {
        // self pair (always shift < 0): order-dependent greedy matching with tombstones
        const int rel = -shift;
        const uint32_t live = P0 | N0;
        uint32_t tomb = 0;
        for (uint32_t m = live; m; m &= m - 1) {
            const int pl = __ffs(m) - 1;
            if ((tomb >> pl) & 1)
                continue;
            const int q = pl + rel;
            if (q >= p.nbits || q >= 32)
                continue;
            if (!((live >> q) & 1) || ((tomb >> q) & 1))
                continue;
            const int s0 = (N0 >> pl) & 1, s1 = (N0 >> q) & 1;
            if ((s0 ^ s1) != sub)
                continue;
            if (s1)
                Nn |= 1u << q;
            else
                Pn |= 1u << q;
            tomb |= (1u << pl) | (1u << q);
        }
        P0 &= ~tomb;
        N0 &= ~tomb;
        P1 = P0;
        N1 = N0;
        pos1 = pos0;
    }
    // placement of the new row: reuse a slot that just died, else any dead slot, else append
    int posn = -1;
    if (Pn | Nn) {
        if ((P0 | N0) == 0 && pos0 >= 0)
            posn = pos0;
        else if (c1 != c0 && (P1 | N1) == 0 && pos1 >= 0)
            posn = pos1;
        else {
            for (int k0 = 0; k0 < len && posn < 0; k0 += 32) {
                const int k = k0 + lane;
                const bool dead = k < len && (L.P[k] | L.N[k]) == 0 && k != pos0 && k != pos1;
                const unsigned m = __ballot_sync(0xffffffffu, dead);
                if (m)
                    posn = k0 + __ffs(m) - 1;
            }
            if (posn < 0) {
                if (len < L.cap)
                    posn = len;
                else if (lane == 0)
                    cx.b->status = ST_LIST_OVERFLOW;
            }
        }
    }
    __syncwarp(); // every lane's reads of the list precede lane 0's in-place update
    if (lane == 0) {
        if (pos0 >= 0 && posn != pos0) {
            L.P[pos0] = P0;
            L.N[pos0] = N0;
        }
        if (pos1 >= 0 && c1 != c0 && posn != pos1) {
            L.P[pos1] = P1;
            L.N[pos1] = N1;
        }
        if (posn >= 0) {
            L.e[posn] = newid;
            L.P[posn] = Pn;
            L.N[posn] = Nn;
            if (posn == len) {
                *L.len = len + 1;
                atomicMax(&cx.b->list_max, len + 1);
            }
        }
        const int a = atomicAdd(&cx.b->n_act, 1);
        ActCol &A = cx.act[a];
        A.o = o;
        A.slot = slot;
        // rows that died and were not recycled keep their slot with empty planes: excluded from the recount by position
        A.pos0 = pos0;
        A.pos1 = (c1 != c0) ? pos1 : pos0;
        A.posn = posn;
        A.P0 = P0, A.N0 = N0, A.P1 = P1, A.N1 = N1, A.Pn = Pn, A.Nn = Nn;
    }
}

// Recount (the column's share of update_stats, state_opr.cc:307-340), executed by the whole CTA over the
// flattened (touched column, row) space.  Every thread builds up to three pair sources (its row against the
// rewritten rows of c0, c1 and the new expression); the warp then walks pair indices in lock step, four L2 atomics
// per lane in flight, their return values inspected afterwards.  Everything stays in registers.
#ifndef DA_RECOUNT_UNROLL
#define DA_RECOUNT_UNROLL 4 // measured: 12 in flight (own register budget via noinline) is slower, the L2 atomic units are the limit
#endif
__device__ void recount_active(const ProblemDesc &p, const Ctx &cx, uint32_t c0, uint32_t c1, uint32_t newid) {
    const int tid = threadIdx.x, nt = blockDim.x;
    BlockCtx &b = *cx.b;
    const int n_act = b.n_act;
    if (n_act == 0)
        return;
    int total = 0;
    for (int a = 0; a < n_act; ++a)
        total += *col_ref(cx, p, cx.act[a].slot, cx.act[a].o).len;
    const int total_pad = (total + 31) & ~31; // whole warps enter the loop together
    int nr = 0;
    for (int item = tid; item < total_pad; item += nt) {
        PairSrc s0, s1, s2;
        s0.n = s1.n = s2.n = 0;
        s0.dhi = s1.dhi = s2.dhi = 1;
        if (item < total) {
            int a = 0, k = item;
            for (;; ++a) {
                const int span = *col_ref(cx, p, cx.act[a].slot, cx.act[a].o).len;
                if (k < span)
                    break;
                k -= span;
            }
            const ActCol &C = cx.act[a];
            const ColRef L = col_ref(cx, p, C.slot, C.o);
            const uint32_t P = L.P[k], N = L.N[k];
            if ((P | N) != 0 && k != C.pos0 && k != C.pos1 && k != C.posn) {
                const uint32_t x = L.e[k];
                const bool h0 = (C.P0 | C.N0) != 0, h1 = (c1 != c0) && ((C.P1 | C.N1) != 0), hn = (C.Pn | C.Nn) != 0;
                s0 = (x < c0) ? make_src(p, h0, 0, x, P, N, C.P0, C.N0) : make_src(p, h0, 0, x, C.P0, C.N0, P, N);
                s1 = (x < c1) ? make_src(p, h1, 1, x, P, N, C.P1, C.N1) : make_src(p, h1, 1, x, C.P1, C.N1, P, N);
                s2 = make_src(p, hn, 2, x, P, N, C.Pn, C.Nn); // x < newid always
            }
        }
        const int n01 = s0.n + s1.n, n_all = n01 + s2.n;
        nr += n_all;
        constexpr int U = DA_RECOUNT_UNROLL; // L2 atomics in flight per lane (their latency is the limiter of the dense early steps)
        for (int base = 0; __any_sync(0xffffffffu, base < n_all); base += U) {
            uint32_t idx[U], old[U];
            bool on[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = base + u;
                on[u] = j < n_all;
                idx[u] = 0u;
                if (on[u])
                    idx[u] = j < s0.n ? pair_index(p, s0, j) : (j < n01 ? pair_index(p, s1, j - s0.n) : pair_index(p, s2, j - n01));
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (on[u])
                    old[u] = atomicAdd(&cx.ws.slab[idx[u]], 1u);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (on[u] && old[u] == 0u)
                    touch_push(cx, idx[u]);
        }
    }
    // pairs among the rewritten rows themselves (dedup rule state_opr.cc:310-312): slot of the larger id
    const int role = nt - 1 - tid;
    if (role < 6) {
        for (int a = 0; a < n_act; ++a) {
            const ActCol &C = cx.act[a];
            const bool h0 = (C.P0 | C.N0) != 0, h1 = (c1 != c0) && ((C.P1 | C.N1) != 0), hn = (C.Pn | C.Nn) != 0;
            if (role == 0 && h0)
                nr += pairs_self(p, cx, 0, c0, C.P0, C.N0);
            if (role == 1 && h1)
                nr += pairs_self(p, cx, 1, c1, C.P1, C.N1);
            if (role == 2 && hn)
                nr += pairs_self(p, cx, 2, newid, C.Pn, C.Nn);
            PairSrc r;
            r.n = 0;
            if (role == 3 && h0 && h1)
                r = make_src(p, true, 1, c0, C.P0, C.N0, C.P1, C.N1); // c0 < c1
            if (role == 4 && h0 && hn)
                r = make_src(p, true, 2, c0, C.P0, C.N0, C.Pn, C.Nn);
            if (role == 5 && h1 && hn)
                r = make_src(p, true, 2, c1, C.P1, C.N1, C.Pn, C.Nn);
            for (int j = 0; j < r.n; ++j) {
                const uint32_t idx = pair_index(p, r, j);
                if (atomicAdd(&cx.ws.slab[idx], 1u) == 0u)
                    touch_push(cx, idx);
            }
            nr += r.n;
        }
    }
    // one 32-bit shared-memory add per warp (64-bit shared atomics are CAS spin loops)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        nr += __shfl_xor_sync(0xffffffffu, nr, off);
    if ((tid & 31) == 0 && nr)
        atomicAdd(&b.r_step, nr);
}

// ---- lazy histogram -------------------------------------------------------------------------------
__device__ __forceinline__ bool entry_live(const FEnt &e, const uint32_t *mod, uint32_t c0, uint32_t c1, bool purge) {
    if (e.y == DA_DEAD)
        return false;
    const uint64_t key = ((uint64_t)e.w << 32) | e.z;
    const uint32_t a = key_id0(key), c = key_id1(key);
    if (purge && (a == c0 || a == c1 || c == c0 || c == c1))
        return false;
    const uint32_t ma = __ldcg(&mod[a]), mc = __ldcg(&mod[c]);
    return e.y >= ma && e.y >= mc;
}

// Re-read one chunk with one warp: rebuild its cached maximum, bury entries found dead, return live count.
// Loads are issued in batches (8 entries per lane, then their 16 stamp lookups) so the round trips overlap.
__device__ __noinline__ int rescan_chunk(const Ctx &cx, int chunk, uint32_t c0, uint32_t c1, bool purge, uint32_t thresh) {
    const int lane = threadIdx.x & 31;
    const int ch = 1 << cx.cfg.chunk_log;
    const int base = chunk << cx.cfg.chunk_log;
    const int end = min(base + ch, cx.b->seg_len);
    const uint32_t *mod = cx.ws.mod_step;
    Best best{0u, 0u, 0u};
    int live = 0;
    constexpr int U = 8;
    for (int i0 = base + lane; i0 < end; i0 += 32 * U) {
        FEnt e[U];
        uint32_t ma[U], mc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + 32 * u;
            e[u] = make_uint4(0u, DA_DEAD, 0u, 0u);
            if (i < end)
                e[u] = __ldcg(&cx.seg[i]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ma[u] = mc[u] = 0u;
            if (e[u].y != DA_DEAD) {
                const uint64_t key = ((uint64_t)e[u].w << 32) | e[u].z;
                ma[u] = __ldcg(&mod[key_id0(key)]);
                mc[u] = __ldcg(&mod[key_id1(key)]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e[u].y == DA_DEAD)
                continue;
            const uint64_t key = ((uint64_t)e[u].w << 32) | e[u].z;
            const uint32_t a = key_id0(key), c = key_id1(key);
            const bool ok = !(purge && (a == c0 || a == c1 || c == c0 || c == c1)) && e[u].y >= ma[u] && e[u].y >= mc[u];
            if (ok) {
                ++live;
                if (e[u].x >= thresh) {
                    Best cand{e[u].x, e[u].w, e[u].z};
                    if (best_gt(cand, best))
                        best = cand;
                }
            }
            else
                cx.seg[i0 + 32 * u].y = DA_DEAD;
        }
    }
    best = warp_best(best);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        live += __shfl_xor_sync(0xffffffffu, live, off);
    if (lane == 0) {
        cx.cb_score[chunk] = best.score;
        cx.cb_khi[chunk] = best.khi;
        cx.cb_klo[chunk] = best.klo;
        cx.cb_dirty[chunk] = 0;
    }
    return live;
}

// Bring every chunk cache up to date for the substitution (c0, c1): chunks whose cached winner touches c0/c1,
// chunks that received appends, or all chunks (accounting / after compaction).  Block-wide.
__device__ void refresh_chunks(const Ctx &cx, uint32_t c0, uint32_t c1, bool purge, bool all, uint32_t thresh) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    BlockCtx &b = *cx.b;
    const int nchunks = (b.seg_len + (1 << cx.cfg.chunk_log) - 1) >> cx.cfg.chunk_log;
    for (int c = tid; c < nchunks; c += nt) {
        bool d = all || cx.cb_dirty[c];
        if (!d && purge && cx.cb_score[c] != 0u) {
            const uint64_t key = ((uint64_t)cx.cb_khi[c] << 32) | cx.cb_klo[c];
            const uint32_t a = key_id0(key), e = key_id1(key);
            d = (a == c0 || a == c1 || e == c0 || e == c1);
        }
        if (d)
            cx.dirty_list[atomicAdd(&b.n_dirty, 1)] = c;
    }
    __syncthreads();
    const int nd = b.n_dirty;
    int live = 0;
    for (int i = wid; i < nd; i += nw)
        live += rescan_chunk(cx, cx.dirty_list[i], c0, c1, purge, thresh);
    if (lane == 0) {
        if (all && live)
            atomicAdd(&b.live_old, live);
        if (nd > wid) {
            const int mine = (nd - wid + nw - 1) / nw;
            atomicAdd(&b.rescan_step, mine << cx.cfg.chunk_log);
        }
    }
    __syncthreads();
    if (tid == 0)
        b.n_dirty = 0;
}

// In-place compaction of this CTA's segment (drops dead entries; order is irrelevant), then every chunk cache
// is rebuilt.  Tiles of 8 x blockDim entries: all reads of a tile complete before its survivors are written to
// positions that never pass the tile's end.
__device__ __noinline__ void compact_segment(const Ctx &cx, uint32_t c0, uint32_t c1, bool purge, uint32_t thresh, long long *compactions) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    BlockCtx &b = *cx.b;
    const int len = b.seg_len;
    constexpr int K = 8;
    if (tid == 0)
        b.cmp_out = 0;
    __syncthreads();
    for (int base = 0; base < len; base += K * nt) {
        FEnt e[K];
        bool live[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = base + k * nt + tid;
            live[k] = false;
            if (i < len) {
                e[k] = __ldcg(&cx.seg[i]);
                live[k] = entry_live(e[k], cx.ws.mod_step, c0, c1, purge);
            }
        }
        __syncthreads(); // every read of this tile is complete
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned bal = __ballot_sync(0xffffffffu, live[k]);
            int wbase = 0;
            if (lane == 0 && bal)
                wbase = atomicAdd(&b.cmp_out, __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (live[k])
                cx.seg[wbase + __popc(bal & ((1u << lane) - 1u))] = e[k];
        }
    }
    __syncthreads();
    if (tid == 0) {
        b.seg_len = b.cmp_out;
        b.live_old = 0;
        if (compactions)
            atomicAdd((unsigned long long *)compactions, 1ull);
    }
    // every cached maximum referred to the old positions: forget them all (the chunks in use are rebuilt below)
    for (int c = tid; c < cx.cfg.nchunk_cap; c += nt) {
        cx.cb_score[c] = 0u;
        cx.cb_khi[c] = 0u;
        cx.cb_klo[c] = 0u;
        cx.cb_dirty[c] = 0;
    }
    __syncthreads();
    refresh_chunks(cx, c0, c1, false, true, thresh);
    if (tid == 0 && !cx.cfg.accounting)
        b.live_old = 0;
    __syncthreads();
}

// Block-reduce (thread bests + chunk caches) and publish this CTA's candidate through the exchange.
__device__ void publish_best(const Ctx &cx, Best best) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    BlockCtx &b = *cx.b;
    const int nchunks = (b.seg_len + (1 << cx.cfg.chunk_log) - 1) >> cx.cfg.chunk_log;
    for (int c = tid; c < nchunks; c += blockDim.x) {
        Best cand{cx.cb_score[c], cx.cb_khi[c], cx.cb_klo[c]};
        if (best_gt(cand, best))
            best = cand;
    }
    best = warp_best(best);
    if (lane == 0)
        b.warp_best[wid] = best;
    __syncthreads();
    if (wid == 0) {
        Best v = lane < nw ? b.warp_best[lane] : Best{0u, 0u, 0u};
        v = warp_best(v);
        if (lane == 0) {
            const unsigned long long key = ((unsigned long long)v.khi << 32) | v.klo;
            const unsigned long long live = (unsigned long long)(cx.cfg.accounting ? b.live_old + b.n_new : 0) & 0x0fffffffULL;
            const unsigned long long want = (b.seg_len > cx.ws.fseg_cap - (cx.ws.fseg_cap >> 2)) ? 1ULL : 0ULL; // ask the whole group to compact together
            xchg_publish(cx, (unsigned long long)v.score | ((unsigned long long)b.status << 32) | (want << 36), key >> 16, (key & 0xffffULL) | (live << 16));
        }
    }
}

// Gather every CTA's candidate -> chosen pair (identical on every CTA); returns |F| (accounting mode).
__device__ int collect_best(const Ctx &cx) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    BlockCtx &b = *cx.b;
    xchg_collect(cx);
    Best v{0u, 0u, 0u};
    int live = 0, st = 0;
    for (int i = tid; i < cx.cfg.G; i += blockDim.x) {
        const unsigned long long w0 = b.xw0[i], w1 = b.xw1[i], w2 = b.xw2[i];
        if ((w0 >> 36) & 1ULL)
            st |= 0x100; // somebody's segment is filling up
        const unsigned long long key = (w1 << 16) | (w2 & 0xffffULL);
        Best c{(uint32_t)w0, (uint32_t)(key >> 32), (uint32_t)key};
        if (best_gt(c, v))
            v = c;
        live += (int)((w2 >> 16) & 0x0fffffffULL);
        st = max(st & 0xff, (int)((w0 >> 32) & 0xf)) | (st & 0x100);
    }
    v = warp_best(v);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        live += __shfl_xor_sync(0xffffffffu, live, off);
        const int o = __shfl_xor_sync(0xffffffffu, st, off);
        st = max(st & 0xff, o & 0xff) | ((st | o) & 0x100);
    }
    if (lane == 0) {
        b.warp_best[wid] = v;
        b.warp_sum[wid] = live;
        b.warp_st[wid] = st;
    }
    __syncthreads();
    if (wid == 0) {
        Best w = lane < nw ? b.warp_best[lane] : Best{0u, 0u, 0u};
        int l = lane < nw ? b.warp_sum[lane] : 0;
        int s2 = lane < nw ? b.warp_st[lane] : 0;
        w = warp_best(w);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            l += __shfl_xor_sync(0xffffffffu, l, off);
            const int o = __shfl_xor_sync(0xffffffffu, s2, off);
            s2 = max(s2 & 0xff, o & 0xff) | ((s2 | o) & 0x100);
        }
        if (lane == 0) {
            b.chosen = w;
            b.scratch_i[0] = l;
            b.scratch_i[1] = s2 & 0xff;
            b.scratch_i[3] = (s2 >> 8) & 1;
        }
    }
    __syncthreads();
    return b.scratch_i[0];
}

// Gather every CTA's touched-counter count; builds the exclusive prefix b.xprefix[0..G]; returns max status.
__device__ int collect_touch_counts(const Ctx &cx) {
    BlockCtx &b = *cx.b;
    xchg_collect(cx);
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int carry = 0, st = 0;
        for (int i0 = 0; i0 < cx.cfg.G; i0 += 32) {
            const int i = i0 + lane;
            const unsigned long long w0 = i < cx.cfg.G ? b.xw0[i] : 0ULL;
            int v = (int)(uint32_t)w0;
            st = max(st, (int)((w0 >> 32) & 0xf));
            int incl = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off)
                    incl += o;
            }
            if (i < cx.cfg.G)
                b.xprefix[i] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            st = max(st, __shfl_xor_sync(0xffffffffu, st, off));
        if (lane == 0) {
            b.xprefix[cx.cfg.G] = carry;
            b.scratch_i[2] = st;
        }
    }
    __syncthreads();
    return b.scratch_i[2];
}

// ---- to_solution (cmvm_core.cc:89-225): one warp per output column ---------------------------------
struct HeapEnt {
    float lat, qmin, qmax, qstep;
    int sub;
    long long la;
    int id, shift;
};
__device__ __forceinline__ bool heap_less(const HeapEnt &a, const HeapEnt &b) {
    // std::tuple operator< over (lat, sub, left_align, qmin, qmax, qstep, id, shift)
    if (a.lat < b.lat)
        return true;
    if (b.lat < a.lat)
        return false;
    if (a.sub != b.sub)
        return a.sub < b.sub;
    if (a.la != b.la)
        return a.la < b.la;
    if (a.qmin < b.qmin)
        return true;
    if (b.qmin < a.qmin)
        return false;
    if (a.qmax < b.qmax)
        return true;
    if (b.qmax < a.qmax)
        return false;
    if (a.qstep < b.qstep)
        return true;
    if (b.qstep < a.qstep)
        return false;
    if (a.id != b.id)
        return a.id < b.id;
    return a.shift < b.shift;
}
__device__ __forceinline__ void heap_store(uint4 *h, int k, const HeapEnt &e) {
    h[2 * k] = make_uint4(__float_as_uint(e.lat), __float_as_uint(e.qmin), __float_as_uint(e.qmax), __float_as_uint(e.qstep));
    h[2 * k + 1] = make_uint4((uint32_t)e.sub | ((uint32_t)e.shift << 8), (uint32_t)e.id, (uint32_t)(unsigned long long)e.la, (uint32_t)((unsigned long long)e.la >> 32));
}
__device__ __forceinline__ HeapEnt heap_load(const uint4 *h, int k) {
    uint4 a = h[2 * k], c = h[2 * k + 1];
    HeapEnt e;
    e.lat = __uint_as_float(a.x);
    e.qmin = __uint_as_float(a.y);
    e.qmax = __uint_as_float(a.z);
    e.qstep = __uint_as_float(a.w);
    e.sub = (int)(c.x & 0xff);
    e.shift = (int)(c.x >> 8);
    e.id = (int)c.y;
    e.la = (long long)(((unsigned long long)c.w << 32) | c.z);
    return e;
}
__device__ __forceinline__ long long left_align(const QInt &q, int shift) {
    float x = fmaxf_std(fabsf(fadd(q.max, q.step)), fabsf(q.min));
    long long n_int = trunc_i64(log2f_ref(x));
    return n_int + (long long)shift; // n_int == INT64_MIN only for degenerate intervals; shift >= 0
}
// Lane-private heap storage: every lane keeps its own entries in its own slice of the scratch
// arena, entries only ever cross lanes through shuffles, so no lane reads memory another lane wrote.
// Remove and return the global minimum over all lanes' private lists (hl[0..cnt) per lane).
__device__ HeapEnt heap_pop(uint4 *hl, int &cnt) {
    const int lane = threadIdx.x & 31;
    HeapEnt best;
    best.lat = 0.0f, best.qmin = 0.0f, best.qmax = 0.0f, best.qstep = 0.0f, best.sub = 0, best.la = 0, best.id = 0, best.shift = 0;
    int bi = -1;
    for (int k = 0; k < cnt; ++k) {
        HeapEnt e = heap_load(hl, k);
        if (bi < 0 || heap_less(e, best)) {
            best = e;
            bi = k;
        }
    }
    // tournament over lanes: winner lane id travels with the candidate
    HeapEnt w = best;
    int wl = bi >= 0 ? lane : -1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        HeapEnt o;
        o.lat = __shfl_xor_sync(0xffffffffu, w.lat, off);
        o.qmin = __shfl_xor_sync(0xffffffffu, w.qmin, off);
        o.qmax = __shfl_xor_sync(0xffffffffu, w.qmax, off);
        o.qstep = __shfl_xor_sync(0xffffffffu, w.qstep, off);
        o.sub = __shfl_xor_sync(0xffffffffu, w.sub, off);
        o.la = __shfl_xor_sync(0xffffffffu, w.la, off);
        o.id = __shfl_xor_sync(0xffffffffu, w.id, off);
        o.shift = __shfl_xor_sync(0xffffffffu, w.shift, off);
        const int ol = __shfl_xor_sync(0xffffffffu, wl, off);
        if (ol >= 0 && (wl < 0 || heap_less(o, w))) {
            w = o;
            wl = ol;
        }
    }
    // (id, shift) is unique per entry, so the order is total and every lane holds the same winner
    if (lane == wl) {
        if (bi != cnt - 1) {
            hl[2 * bi] = hl[2 * (cnt - 1)];
            hl[2 * bi + 1] = hl[2 * (cnt - 1) + 1];
        }
        cnt -= 1;
    }
    return w;
}

__device__ __noinline__ void column_finish(const ProblemDesc &p, const Ctx &cx, int slot, int o, int gid_base) {
    const int lane = threadIdx.x & 31;
    const ColRef L = col_ref(cx, p, slot, o);
    const int len = *L.len;
    uint4 *hl = cx.ws.heap + 2 * ((size_t)o * 32 + lane) * (size_t)p.heap_lane_cap;
    // digits of the rows k = lane (mod 32) go to this lane's private list
    int cnt = 0;
    for (int k = lane; k < len; k += 32) {
        const uint32_t ce = L.e[k], cP = L.P[k], cN = L.N[k];
        for (uint32_t m = cP | cN; m; m &= m - 1) {
            const int sh = __ffs(m) - 1;
            QInt q;
            float lat;
            load_op(p, ce, q, lat);
            HeapEnt e;
            e.lat = lat;
            e.sub = (int)((cN >> sh) & 1);
            e.la = left_align(q, sh);
            e.qmin = q.min;
            e.qmax = q.max;
            e.qstep = q.step;
            e.id = (int)ce;
            e.shift = sh;
            if (cnt < p.heap_lane_cap)
                heap_store(hl, cnt, e);
            ++cnt;
        }
    }
    cnt = min(cnt, p.heap_lane_cap);
    int n = cnt;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        n += __shfl_xor_sync(0xffffffffu, n, off);
    const int base_shift = (int)p.shift1[o];
    if (n == 0) {
        if (lane == 0) {
            p.out_idx[o] = -1;
            p.out_shift[o] = base_shift;
            p.out_neg[o] = 0;
            p.out_q[o] = make_float4(0.0f, 0.0f, __uint_as_float(0x7f800000u), 0.0f); // api.cc:110-113
        }
        return;
    }
    int gid = gid_base;
    while (n > 1) {
        const HeapEnt e0 = heap_pop(hl, cnt);
        const HeapEnt e1 = heap_pop(hl, cnt);
        // every lane holds (e0, e1): compute the merged entry redundantly, lane 0 records the op
        const QInt q0{e0.qmin, e0.qmax, e0.qstep}, q1{e1.qmin, e1.qmax, e1.qstep};
        QInt q;
        float dlat, dcost;
        int4 misc;
        int rshift;
        if (e0.sub) {
            const long long s = (long long)e0.shift - e1.shift;
            q = qint_add(q1, q0, s, e1.sub != 0, e0.sub != 0);
            cost_add(q1, q0, s, (1 ^ e1.sub) != 0, p.adder_size, p.carry_size, dlat, dcost);
            misc = make_int4(e1.id, e0.id, 1 ^ e1.sub, (int)s);
            rshift = e1.shift;
        }
        else {
            const long long s = (long long)e1.shift - e0.shift;
            q = qint_add(q0, q1, s, e0.sub != 0, e1.sub != 0);
            cost_add(q0, q1, s, e1.sub != 0, p.adder_size, p.carry_size, dlat, dcost);
            misc = make_int4(e0.id, e1.id, e1.sub, (int)s);
            rshift = e0.shift;
        }
        const float lat = fadd(fmaxf_std(e0.lat, e1.lat), dlat);
        if (lane == 0 && gid < p.ops_cap) {
            p.op_misc[gid] = misc;
            p.op_q[gid] = make_float4(q.min, q.max, q.step, lat);
            p.op_cost[gid] = dcost;
        }
        if (lane == (gid & 31)) {
            HeapEnt ne;
            ne.lat = lat;
            ne.sub = e0.sub & e1.sub;
            ne.la = left_align(q, rshift);
            ne.qmin = q.min;
            ne.qmax = q.max;
            ne.qstep = q.step;
            ne.id = gid;
            ne.shift = rshift;
            if (cnt < p.heap_lane_cap) {
                heap_store(hl, cnt, ne);
                ++cnt;
            }
        }
        n -= 1;
        gid += 1;
    }
    // the single remaining entry lives in exactly one lane
    if (cnt == 1) {
        const HeapEnt e = heap_load(hl, 0);
        p.out_idx[o] = e.id;
        p.out_neg[o] = e.sub;
        p.out_shift[o] = base_shift + e.shift;
        p.out_q[o] = make_float4(e.qmin, e.qmax, e.qstep, e.lat); // the RAW op interval/latency (api.cc:103-109)
    }
}

// ------------------------------------------------------------------------------------------------
// solve one problem with the CTAs of one group

__device__ void solve_problem(const ProblemDesc &p, const Ctx &cx) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int n_in = p.n_in, n_out = p.n_out, nbits = p.nbits, G = cx.cfg.G;
    const uint32_t thresh = method_threshold(p.method);
    BlockCtx &b = *cx.b;

    if (tid == 0) {
        b.seg_len = 0;
        b.n_new = 0;
        b.live_old = 0;
        b.touch_n = 0;
        b.n_act = 0;
        b.n_dirty = 0;
        b.status = ST_OK;
        b.list_max = 0;
        b.r_count = 0ull;
        b.rescanned = 0ull;
        b.r_step = 0;
        b.rescan_step = 0;
        b.chosen = Best{0u, 0u, 0u};
        for (int k = 0; k < 8; ++k)
            b.phase[k] = 0;
        b.poll_iters = 0;
        for (int k = 0; k < 8; ++k) {
            b.peak[k] = 0;
            b.nslow[k] = 0;
        }
    }
    for (int c = tid; c < cx.cfg.nchunk_cap; c += nt) {
        cx.cb_score[c] = 0u;
        cx.cb_khi[c] = 0u;
        cx.cb_klo[c] = 0u;
        cx.cb_dirty[c] = 0;
    }
    // rewrite stamps start at zero (initial entries carry stamp 0)
    for (int i = cx.rank * nt + tid; i < p.e_cap; i += G * nt)
        cx.ws.mod_step[i] = 0u;
    __syncthreads();

    // ---- column lists (state_opr.cc:100-112): warp per owned column
    for (int slot = wid; slot < cx.cfg.cpc; slot += nw) {
        const int oc = cx.rank + G * slot;
        if (oc >= n_out)
            break;
        const ColRef L = col_ref(cx, p, slot, oc);
        // slot i holds input i in every column (empty planes when the entry is zero): rows of one expression line up
        // across columns, which the recount exploits; empty slots are recycled by later rows
        const int len = n_in;
        for (int i = lane; i < n_in && i < L.cap; i += 32) {
            const uint2 m = p.masks0[(size_t)i * n_out + oc];
            L.e[i] = (uint32_t)i;
            L.P[i] = m.x;
            L.N[i] = m.y;
        }
        if (lane == 0) {
            if (len > L.cap)
                b.status = ST_LIST_OVERFLOW;
            *L.len = min(len, L.cap);
            atomicMax(&b.list_max, min(len, L.cap));
        }
    }
    // ---- input ops (state_opr.cc:146-149)
    for (int i = cx.rank * nt + tid; i < n_in; i += G * nt) {
        p.op_misc[i] = make_int4(i, -1, -1, 0);
        p.op_q[i] = make_float4(p.qint[3 * i], p.qint[3 * i + 1], p.qint[3 * i + 2], p.lat[i]);
        p.op_cost[i] = 0.0f;
    }

    Best best{0u, 0u, 0u};
    unsigned long long r0 = 0;
    if (p.method != M_DUMMY) {
        // ---- initial pair histogram (state_opr.cc:115-144, types.hh:73-100): warp per (a<=c) pair block,
        //      lanes over relative shifts, sign planes streamed over the output columns
        const long long n_pairs = (long long)n_in * (n_in + 1) / 2;
        const int n_sh = 2 * nbits - 1;
        for (long long pi = (long long)cx.rank * nw + wid; pi < n_pairs; pi += (long long)G * nw) {
            long long a = (long long)(((2.0 * n_in + 1.0) - sqrt((2.0 * n_in + 1.0) * (2.0 * n_in + 1.0) - 8.0 * (double)pi)) * 0.5);
            while (a > 0 && a * (2LL * n_in - a + 1) / 2 > pi)
                --a;
            while ((a + 1) * (2LL * n_in - (a + 1) + 1) / 2 <= pi)
                ++a;
            const long long c = a + (pi - a * (2LL * n_in - a + 1) / 2);
            const uint2 *ra = p.masks0 + (size_t)a * n_out;
            const uint2 *rc = p.masks0 + (size_t)c * n_out;
            QInt qa, qc;
            float la, lc;
            load_op(p, (uint32_t)a, qa, la);
            load_op(p, (uint32_t)c, qc, lc);
            for (int s0 = 0; s0 < n_sh; s0 += 32) {
                const int si = s0 + lane;
                const int s = si - (nbits - 1);
                const bool active = si < n_sh && !(a == c && s >= 0);
                uint32_t same = 0, diff = 0;
                if (active) {
                    if (s >= 0) {
                        for (int o = 0; o < n_out; ++o) {
                            const uint2 x = ra[o], y = rc[o];
                            same += __popc(x.x & (y.x >> s)) + __popc(x.y & (y.y >> s));
                            diff += __popc(x.x & (y.y >> s)) + __popc(x.y & (y.x >> s));
                        }
                    }
                    else {
                        const int d = -s;
                        for (int o = 0; o < n_out; ++o) {
                            const uint2 x = ra[o], y = rc[o];
                            same += __popc((x.x >> d) & y.x) + __popc((x.y >> d) & y.y);
                            diff += __popc((x.x >> d) & y.y) + __popc((x.y >> d) & y.x);
                        }
                    }
                    r0 += same + diff;
                    if (same >= 2)
                        emit_entry(p, cx, (uint32_t)a, (uint32_t)c, s, 0, same, qa, la, qc, lc, 0u, thresh, best);
                    if (diff >= 2)
                        emit_entry(p, cx, (uint32_t)a, (uint32_t)c, s, 1, diff, qa, la, qc, lc, 0u, thresh, best);
                }
            }
        }
    }
    if (r0)
        atomicAdd(&b.r_count, r0);
    __syncthreads();
    const unsigned long long r0_cta = b.r_count;
    if (tid == 0) {
        b.seg_len = min(b.seg_len, cx.ws.fseg_cap);
        b.n_new = 0;
    }
    __syncthreads();
    // build every chunk cache (and the exact initial size)
    refresh_chunks(cx, 0u, 0u, false, true, thresh);
    publish_best(cx, Best{0u, 0u, 0u}); // the exchange is also the barrier that orders mod_step zeroing / list building
    int f_live = collect_best(cx);
    const int f0 = f_live;
    int f_max = f_live;

    // ---- greedy loop (cmvm_core.cc:36-70)
    int t = 0;
    unsigned long long sum_f = 0;
    int status = b.scratch_i[1];
    while (status == ST_OK) {
        const Best ch = b.chosen;
        if (ch.score == 0u || p.method == M_DUMMY)
            break;
        if (cx.cfg.max_steps > 0 && t >= cx.cfg.max_steps)
            break;
        if (n_in + t >= p.e_cap) {
            status = ST_EXPR_OVERFLOW;
            break;
        }
        const uint64_t key = ((uint64_t)ch.khi << 32) | ch.klo;
        const uint32_t c0 = key_id0(key), c1 = key_id1(key);
        const int shift = key_shift(key), sub = key_sub(key);
        const uint32_t newid = (uint32_t)(n_in + t);
        const uint32_t stamp = (uint32_t)(t + 1);
        sum_f += (unsigned long long)f_live;
        f_max = max(f_max, f_live);
        if (cx.rank == 0 && tid == nt - 1) {
            // pair_to_op (state_opr.cc:211-225) + rewrite stamps; published by the next group barrier
            QInt q0, q1;
            float l0, l1;
            load_op(p, c0, q0, l0);
            load_op(p, c1, q1, l1);
            float dlat, cost;
            cost_add(q0, q1, shift, sub != 0, p.adder_size, p.carry_size, dlat, cost);
            const QInt q = qint_add(q0, q1, shift, false, sub != 0);
            const float lat = fadd(fmaxf_std(l0, l1), dlat);
            p.op_misc[newid] = make_int4((int)c0, (int)c1, sub, shift);
            p.op_q[newid] = make_float4(q.min, q.max, q.step, lat);
            p.op_cost[newid] = cost;
            cx.ws.mod_step[c0] = stamp;
            cx.ws.mod_step[c1] = stamp;
            cx.ws.mod_step[newid] = stamp;
            if (p.trace && t < p.trace_cap) {
                int *tr = p.trace + 5 * (size_t)t;
                tr[0] = (int)c0;
                tr[1] = (int)c1;
                tr[2] = shift;
                tr[3] = sub;
                tr[4] = f_live;
            }
        }
        if (tid == 0)
            b.t_last = clock64();
        // A. substitute in the owned columns (one warp each), then recount with the whole CTA
        for (int slot = wid; slot < cx.cfg.cpc; slot += nw) {
            const int oc = cx.rank + G * slot;
            if (oc >= n_out)
                break;
            column_substitute(p, cx, slot, oc, c0, c1, shift, sub, newid);
        }
        __syncthreads();
        DA_LAP(0)
        recount_active(p, cx, c0, c1, newid);
        __syncthreads();
        if (tid == 0) {
            // exchange 1: this CTA's touched-counter count (the fence inside also publishes the op record / stamps)
            xchg_publish(cx, (unsigned long long)(uint32_t)min(b.touch_n, cx.ws.touch_cap) | ((unsigned long long)b.status << 32), 0ULL, 0ULL);
            b.n_act = 0;
            b.live_old = 0;
            b.n_new = 0;
            b.touch_n = 0;
            b.r_count += (unsigned long long)b.r_step;
            b.r_step = 0;
            b.rescanned += (unsigned long long)b.rescan_step;
            b.rescan_step = 0;
        }
        DA_LAP(1)
        __syncthreads();
        // B. refresh the argmax caches while the other CTAs finish their columns
        if (b.scratch_i[3]) // agreed by the whole group in the last exchange: everybody compacts in the same step
            compact_segment(cx, c0, c1, true, thresh, cx.rank == 0 ? &p.result_meta[META_COMPACTIONS] : nullptr);
        else
            refresh_chunks(cx, c0, c1, true, cx.cfg.accounting != 0, thresh);
        DA_LAP(2)
        const int st1 = collect_touch_counts(cx);
        DA_LAP(3)
        if (st1 != ST_OK) { // identical on every CTA
            status = st1;
            break;
        }
        // C. harvest an equal share of all touched counters of the group (balanced, whoever touched them)
        const int n_all = b.xprefix[G];
        const int h_lo = (int)((long long)n_all * cx.rank / G), h_hi = (int)((long long)n_all * (cx.rank + 1) / G);
        const int n_mine = h_hi - h_lo;
        if (b.seg_len + n_mine > cx.ws.fseg_cap) // (uniform over the CTA)
            compact_segment(cx, c0, c1, true, thresh, cx.rank == 0 ? &p.result_meta[META_COMPACTIONS] : nullptr);
        best = Best{0u, 0u, 0u};
        for (int i = tid; i < n_mine; i += nt) {
            const int gidx = h_lo + i;
            int lo_s = 0, hi_s = G; // largest src with xprefix[src] <= gidx
            while (hi_s - lo_s > 1) {
                const int mid = (lo_s + hi_s) >> 1;
                if (b.xprefix[mid] <= gidx)
                    lo_s = mid;
                else
                    hi_s = mid;
            }
            const uint32_t idx = __ldcg(&cx.ws.touch[(size_t)lo_s * cx.ws.touch_cap + (gidx - b.xprefix[lo_s])]);
            const int sb = (int)(idx & 1u);
            const int si = (int)((idx >> 1) & ((1u << (p.log_s - 1)) - 1u));
            const uint32_t r = idx >> p.log_s;
            const int slot = r >= 2u * (uint32_t)p.e_cap ? 2 : (r >= (uint32_t)p.e_cap ? 1 : 0);
            const uint32_t x = r - (uint32_t)slot * (uint32_t)p.e_cap;
            const uint32_t m = slot == 0 ? c0 : (slot == 1 ? c1 : newid);
            const uint32_t lo = min(m, x), hi = max(m, x);
            // counter and both operand records are fetched together (one L2 round trip)
            const uint32_t cnt = __ldcg(&cx.ws.slab[idx]);
            QInt q0, q1;
            float l0, l1;
            load_op(p, lo, q0, l0);
            load_op(p, hi, q1, l1);
            cx.ws.slab[idx] = 0u;
            if (cnt >= 2)
                emit_entry(p, cx, lo, hi, si - (nbits - 1), sb, cnt, q0, l0, q1, l1, stamp, thresh, best);
        }
        __syncthreads();
        DA_LAP(4)
        publish_best(cx, best);
        DA_LAP(5)
        f_live = collect_best(cx);
        DA_LAP(7)
        ++t;
        if (b.scratch_i[1] != ST_OK) { // some CTA overflowed a buffer: every CTA sees it in the exchange and stops
            status = b.scratch_i[1];
            break;
        }
    }

    // ---- to_solution
    for (int slot = wid; slot < cx.cfg.cpc; slot += nw) {
        const int oc = cx.rank + G * slot;
        if (oc >= n_out)
            break;
        const ColRef L = col_ref(cx, p, slot, oc);
        const int len = *L.len;
        int k = 0;
        for (int i = lane; i < len; i += 32)
            k += __popc(L.P[i]) + __popc(L.N[i]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            k += __shfl_xor_sync(0xffffffffu, k, off);
        if (lane == 0)
            __stcg(&cx.ws.col_k[oc], k);
    }
    group_sync(cx);
    for (int slot = wid; slot < cx.cfg.cpc; slot += nw) {
        const int oc = cx.rank + G * slot;
        if (oc >= n_out)
            break;
        int before = 0;
        for (int i = lane; i < oc; i += 32) {
            const int k = __ldcg(&cx.ws.col_k[i]);
            before += k > 1 ? k - 1 : 0;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            before += __shfl_xor_sync(0xffffffffu, before, off);
        column_finish(p, cx, slot, oc, n_in + t + before);
    }
    group_sync(cx); // every column's tree ops are written
    if (cx.rank == 0 && wid == 0) {
        // float cost of the stage, summed in op order like the reference (api.cc:222-227): warp-wide loads, the
        // additions themselves stay strictly sequential
        long long n_ops_all = (long long)n_in + t;
        for (int o = 0; o < n_out; ++o) {
            const int k = __ldcg(&cx.ws.col_k[o]);
            n_ops_all += k > 1 ? k - 1 : 0;
        }
        n_ops_all = min(n_ops_all, (long long)p.ops_cap);
        float c = p.cost_init;
        for (long long base = 0; base < n_ops_all; base += 32) {
            const float v = base + lane < n_ops_all ? __ldcg(&p.op_cost[base + lane]) : 0.0f;
            const int m = (int)min(32LL, n_ops_all - base);
            for (int k = 0; k < m; ++k)
                c = fadd(c, __shfl_sync(0xffffffffu, v, k));
        }
        if (lane == 0)
            p.result_meta[META_COST_BITS] = (long long)__float_as_uint(c);
    }
    // ---- bookkeeping
    __syncthreads();
    if (tid == 0) {
        b.r_count += (unsigned long long)b.r_step;
        b.rescanned += (unsigned long long)b.rescan_step;
        atomicAdd((unsigned long long *)&p.result_meta[META_SUM_R], b.r_count - r0_cta);
        atomicAdd((unsigned long long *)&p.result_meta[META_R0], r0_cta);
        atomicAdd((unsigned long long *)&p.result_meta[META_RESCANNED], b.rescanned);
        atomicMax((long long *)&p.result_meta[META_LIST_MAX], (long long)b.list_max);
        for (int k = 0; k < 8; ++k)
            atomicMax((long long *)&p.result_meta[META_PHASEMAX + k], b.peak[k] * 1000000LL + b.nslow[k]);
        if (b.status != ST_OK)
            atomicMax((int *)&p.result_meta[META_STATUS], b.status);
    }
    if (cx.rank == 0 && tid == 0) {
        long long tree = 0, dfin = 0;
        for (int o = 0; o < n_out; ++o) {
            const int k = __ldcg(&cx.ws.col_k[o]);
            tree += k > 1 ? k - 1 : 0;
            dfin += k;
        }
        const long long n_ops = (long long)n_in + t + tree;
        p.result_meta[META_N_OPS] = n_ops;
        p.result_meta[META_T] = t;
        p.result_meta[META_SUM_F] = (long long)sum_f;
        p.result_meta[META_F0] = f0;
        p.result_meta[META_D_FINAL] = dfin;
        p.result_meta[META_F_MAX] = f_max;
        for (int k = 0; k < 8; ++k)
            p.result_meta[META_PHASE0 + k] = b.phase[k];
        p.result_meta[15] = b.poll_iters;
        if (status != ST_OK)
            atomicMax((int *)&p.result_meta[META_STATUS], status);
        if (n_ops > p.ops_cap)
            atomicMax((int *)&p.result_meta[META_STATUS], (int)ST_OPS_OVERFLOW);
    }
    group_sync(cx); // the workspace may be reused by the next problem of this group
}

// grid = n_groups * G CTAs; group i solves problems i, i + n_groups, ...
__device__ __forceinline__ void solve_kernel_body(const ProblemDesc *probs, int n_probs, const GroupWs *wss, const LaunchCfg &cfg) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ BlockCtx bctx;
    Ctx cx;
    cx.cfg = cfg;
    cx.rank = blockIdx.x % cfg.G;
    const int group = blockIdx.x / cfg.G, n_groups = gridDim.x / cfg.G;
    cx.ws = wss[group];
    cx.seg = cx.ws.fseg + (size_t)cx.rank * cx.ws.fseg_cap;
    cx.touch_g = cx.ws.touch + (size_t)cx.rank * cx.ws.touch_cap;
    cx.b = &bctx;
    unsigned char *sp = smem;
    cx.cb_score = (uint32_t *)sp;
    sp += sizeof(uint32_t) * cfg.nchunk_cap;
    cx.cb_khi = (uint32_t *)sp;
    sp += sizeof(uint32_t) * cfg.nchunk_cap;
    cx.cb_klo = (uint32_t *)sp;
    sp += sizeof(uint32_t) * cfg.nchunk_cap;
    cx.dirty_list = (int *)sp;
    sp += sizeof(int) * cfg.nchunk_cap;
    cx.col_len_s = (int *)sp;
    sp += sizeof(int) * cfg.cpc;
    cx.act = (ActCol *)sp;
    sp += sizeof(ActCol) * cfg.cpc;
    cx.lists_s = (uint32_t *)sp;
    sp += sizeof(uint32_t) * 3 * (size_t)cfg.cpc * cfg.lcap;
    cx.cb_dirty = sp;
    if (threadIdx.x == 0)
    {
        bctx.bar_target = 0u; // the host zeroes the arrive counter and the exchange slots before every launch
        bctx.epoch = 0u;
    }
    __syncthreads();
    for (int pi = group; pi < n_probs; pi += n_groups)
        solve_problem(probs[pi], cx);
}
// one 512-thread CTA per SM (a lone problem: maximum threads per column / segment) ...
__global__ void __launch_bounds__(512, 1) cmvm_solve_kernel(const ProblemDesc *probs, int n_probs, const GroupWs *wss, LaunchCfg cfg) {
    solve_kernel_body(probs, n_probs, wss, cfg);
}
// ... or two 256-thread CTAs per SM (several problems in flight: one CTA's exchange wait overlaps the other's work)
__global__ void __launch_bounds__(256, 2) cmvm_solve_kernel_x2(const ProblemDesc *probs, int n_probs, const GroupWs *wss, LaunchCfg cfg) {
    solve_kernel_body(probs, n_probs, wss, cfg);
}

// Developer micro-benchmark of the group exchange (not part of the product path): `iters` back-to-back
// publish/collect rounds, optionally with `work` dummy global stores per thread before each publish.
__global__ void __launch_bounds__(512, 1) xchg_bench_kernel(GroupWs ws, int G, int iters, int work, unsigned *sink, long long *cycles) {
    __shared__ BlockCtx bctx;
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.cfg.G = G;
    cx.rank = blockIdx.x % G;
    cx.ws = ws;
    cx.b = &bctx;
    if (threadIdx.x == 0) {
        bctx.bar_target = 0u;
        bctx.epoch = 0u;
    }
    __syncthreads();
    const long long t0 = clock64();
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        for (int w = 0; w < work; ++w)
            sink[(size_t)blockIdx.x * 512 * 8 + (size_t)w * 512 + threadIdx.x] = it + w;
        __syncthreads();
        if (threadIdx.x == 0)
            xchg_publish(cx, (unsigned long long)it, 1ULL, 2ULL);
        xchg_collect(cx);
        acc += (unsigned)bctx.xw0[threadIdx.x % G];
    }
    if (threadIdx.x == 0) {
        cycles[blockIdx.x] = clock64() - t0;
        sink[blockIdx.x] = acc;
    }
}

} // namespace da
