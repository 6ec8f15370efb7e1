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
//   * the histogram is an unordered, append-only log of (score, count, packed key) split into one
//     segment per CTA; the reference's "erase entries touching id0/id1" is a tombstone written while
//     the next argmax scan streams the segment; order independence comes from reducing on the
//     composite (score, key), whose order is exactly the reference's "last maximum in sorted order";
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
    int seg_len;       // entries (live + tombstones) in this CTA's histogram segment
    int seg_live;      // live entries
    int touch_n;       // counters first-touched by this CTA in the current step
    int status;        // sticky error
    unsigned long long r_count; // digit pairs enumerated by this CTA (all steps)
    unsigned bar_target;
    int scratch_i[4];
};

struct GroupCtx {
    int G, rank;       // CTAs in the group, this CTA's index
    GroupWs ws;
    FEnt *seg;         // this CTA's histogram segment
    uint32_t *touch;   // this CTA's touched-counter list
};

// Barrier across the G CTAs of a group (monotonic counter, sense-free).  Split in arrive / wait so
// that independent work can overlap the wait.
__device__ __forceinline__ void group_arrive(const GroupCtx &g, BlockCtx &b) {
    __syncthreads();
    if (g.G > 1 && threadIdx.x == 0) {
        __threadfence();
        atomicAdd(g.ws.barrier, 1u);
        b.bar_target += (unsigned)g.G;
    }
}
__device__ __forceinline__ void group_wait(const GroupCtx &g, BlockCtx &b) {
    if (g.G > 1) {
        if (threadIdx.x == 0) {
            unsigned target = b.bar_target;
            while ((int)(ld_acquire_u32(g.ws.barrier) - target) < 0) {
            }
            __threadfence();
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void group_sync(const GroupCtx &g, BlockCtx &b) {
    group_arrive(g, b);
    group_wait(g, b);
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

// Append one histogram entry to this CTA's segment and fold it into the thread's running best.
__device__ __forceinline__ void
emit_entry(const ProblemDesc &p, const GroupCtx &g, BlockCtx &b, uint32_t lo, uint32_t hi, int shift, int sub, uint32_t count, uint32_t thresh, Best &best) {
    QInt q0, q1;
    float l0, l1;
    load_op(p, lo, q0, l0);
    load_op(p, hi, q1, l1);
    uint32_t score;
    if (!pair_score(p.method, count, q0, l0, q1, l1, score))
        return; // NaN score: can never be selected
    uint64_t key = pack_key(lo, hi, shift, sub);
    int pos = atomicAdd(&b.seg_len, 1);
    if (pos >= g.ws.fseg_cap) {
        b.status = ST_FSEG_OVERFLOW;
        return;
    }
    atomicAdd(&b.seg_live, 1);
    FEnt e;
    e.x = score;
    e.y = count;
    e.z = (uint32_t)key;
    e.w = (uint32_t)(key >> 32);
    g.seg[pos] = e;
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
    __shared__ int s_max, s_d0, s_colcap, s_dcolmax;
    if (tid == 0) {
        s_max = 0;
        s_d0 = 0;
        s_colcap = 0;
        s_dcolmax = 0;
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
        atomicMax(&s_colcap, d + rows);
        atomicMax(&s_dcolmax, d);
    }
    __syncthreads();
    if (tid == 0) {
        p.prep_meta[PM_NBITS] = N;
        p.prep_meta[PM_D0] = s_d0;
        p.prep_meta[PM_COLCAP] = s_colcap;
        p.prep_meta[PM_DCOL_MAX] = s_dcolmax;
    }
}

// ------------------------------------------------------------------------------------------------
// solve: pieces

// counter index in the slab
__device__ __forceinline__ uint32_t slab_index(const ProblemDesc &p, int slot, uint32_t x, int shift, int sub) {
    return (((uint32_t)slot * (uint32_t)p.e_cap + x) << p.log_s) + (uint32_t)(((shift + p.nbits - 1) << 1) | sub);
}

__device__ __forceinline__ void
bump(const ProblemDesc &p, const GroupCtx &g, BlockCtx &b, int slot, uint32_t x, int shift, int sub) {
    uint32_t idx = slab_index(p, slot, x, shift, sub);
    uint32_t old = atomicAdd(&g.ws.slab[idx], 1u);
    if (old == 0) {
        int t = atomicAdd(&b.touch_n, 1);
        if (t < g.ws.touch_cap)
            g.touch[t] = idx;
        else
            b.status = ST_TOUCH_OVERFLOW;
    }
}

// all digit pairs between row lo and row hi (lo != hi), state_opr.cc:331-336
__device__ __forceinline__ int
pairs_cross(const ProblemDesc &p, const GroupCtx &g, BlockCtx &b, int slot, uint32_t x, uint32_t Plo, uint32_t Nlo, uint32_t Phi, uint32_t Nhi) {
    int n = 0;
    for (uint32_t ml = Plo | Nlo; ml; ml &= ml - 1) {
        int pl = __ffs(ml) - 1;
        int sl = (Nlo >> pl) & 1;
        for (uint32_t mh = Phi | Nhi; mh; mh &= mh - 1) {
            int ph = __ffs(mh) - 1;
            int sh = (Nhi >> ph) & 1;
            bump(p, g, b, slot, x, ph - pl, sl ^ sh);
            ++n;
        }
    }
    return n;
}
// digit pairs inside one row, state_opr.cc:323-330: v0 = higher digit, v1 = lower -> negative shift
__device__ __forceinline__ int
pairs_self(const ProblemDesc &p, const GroupCtx &g, BlockCtx &b, int slot, uint32_t x, uint32_t P, uint32_t N) {
    int n = 0;
    for (uint32_t ma = P | N; ma; ma &= ma - 1) {
        int pa = __ffs(ma) - 1;
        int sa = (N >> pa) & 1;
        for (uint32_t mb = (P | N) & ((1u << pa) - 1u); mb; mb &= mb - 1) {
            int pb = __ffs(mb) - 1;
            int sb = (N >> pb) & 1;
            bump(p, g, b, slot, x, pb - pa, sa ^ sb);
            ++n;
        }
    }
    return n;
}

// One greedy step inside one output column, executed by one warp (update_expr state_opr.cc:227-283
// followed by the column's share of update_stats state_opr.cc:307-340).
__device__ void column_step(const ProblemDesc &p, const GroupCtx &g, BlockCtx &b, int o, uint32_t c0, uint32_t c1, int shift, int sub, uint32_t newid) {
    const int lane = threadIdx.x & 31;
    ColEnt *list = g.ws.col_ents + (size_t)o * p.col_cap;
    const int L = g.ws.col_len[o];

    // locate the rows of c0 / c1
    int pos0 = -1, pos1 = -1;
    uint32_t P0 = 0, N0 = 0, P1 = 0, N1 = 0;
    for (int k = lane; k < L; k += 32) {
        uint4 v = *reinterpret_cast<const uint4 *>(&list[k]);
        if (v.x == c0) {
            pos0 = k;
            P0 = v.y;
            N0 = v.z;
        }
        if (v.x == c1) {
            pos1 = k;
            P1 = v.y;
            N1 = v.z;
        }
    }
    {
        unsigned m0 = __ballot_sync(0xffffffffu, pos0 >= 0);
        unsigned m1 = __ballot_sync(0xffffffffu, pos1 >= 0);
        int s0 = m0 ? __ffs(m0) - 1 : 0, s1 = m1 ? __ffs(m1) - 1 : 0;
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

    // ---- substitution (uniform across the warp)
    uint32_t Pn = 0, Nn = 0;
    if (c0 != c1) {
        const bool flip = shift < 0;
        const int rel = flip ? -shift : shift;
        uint32_t AP = flip ? P1 : P0, AN = flip ? N1 : N0; // expr0 after the reference's swap
        uint32_t BP = flip ? P0 : P1, BN = flip ? N0 : N1;
        uint32_t M = sub ? ((AP & (BN >> rel)) | (AN & (BP >> rel))) : ((AP & (BP >> rel)) | (AN & (BN >> rel)));
        uint32_t MB = M << rel;
        uint32_t AP2 = AP & ~M, AN2 = AN & ~M, BP2 = BP & ~MB, BN2 = BN & ~MB;
        if (!flip) { // new digit takes position and sign of id0's digit (expr0)
            Pn = AP & M;
            Nn = AN & M;
            P0 = AP2, N0 = AN2, P1 = BP2, N1 = BN2;
        }
        else { // id0 is expr1 after the swap
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
            int pl = __ffs(m) - 1;
            if ((tomb >> pl) & 1)
                continue;
            int q = pl + rel;
            if (q >= p.nbits || q >= 32)
                continue;
            if (!((live >> q) & 1) || ((tomb >> q) & 1))
                continue;
            int s0 = (N0 >> pl) & 1, s1 = (N0 >> q) & 1;
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
    }
    if (lane == 0) {
        if (pos0 >= 0) {
            list[pos0].P = P0;
            list[pos0].N = N0;
        }
        if (pos1 >= 0 && c1 != c0) {
            list[pos1].P = P1;
            list[pos1].N = N1;
        }
        if (Pn | Nn) {
            if (L < p.col_cap) {
                ColEnt ne{newid, Pn, Nn, 0u};
                list[L] = ne;
                g.ws.col_len[o] = L + 1;
            }
            else
                b.status = ST_LIST_OVERFLOW;
        }
    }
    __syncwarp();

    // ---- recount: modified rows (slot 0: c0, slot 1: c1, slot 2: new) against every other live row
    const bool h0 = (P0 | N0) != 0, h1 = (c1 != c0) && ((P1 | N1) != 0), hn = (Pn | Nn) != 0;
    int nr = 0;
    for (int k = lane; k < L; k += 32) {
        if (k == pos0 || k == pos1)
            continue;
        uint4 v = *reinterpret_cast<const uint4 *>(&list[k]);
        if ((v.y | v.z) == 0)
            continue;
        const uint32_t x = v.x;
        if (h0)
            nr += (x < c0) ? pairs_cross(p, g, b, 0, x, v.y, v.z, P0, N0) : pairs_cross(p, g, b, 0, x, P0, N0, v.y, v.z);
        if (h1)
            nr += (x < c1) ? pairs_cross(p, g, b, 1, x, v.y, v.z, P1, N1) : pairs_cross(p, g, b, 1, x, P1, N1, v.y, v.z);
        if (hn) // x < newid always
            nr += pairs_cross(p, g, b, 2, x, v.y, v.z, Pn, Nn);
    }
    // pairs among the modified rows themselves (dedup rule state_opr.cc:310-312): slot of the larger id
    if (lane == 0 && h0)
        nr += pairs_self(p, g, b, 0, c0, P0, N0);
    if (lane == 1 && h1)
        nr += pairs_self(p, g, b, 1, c1, P1, N1);
    if (lane == 2 && hn)
        nr += pairs_self(p, g, b, 2, newid, Pn, Nn);
    if (lane == 3 && h0 && h1)
        nr += pairs_cross(p, g, b, 1, c0, P0, N0, P1, N1); // c0 < c1
    if (lane == 4 && h0 && hn)
        nr += pairs_cross(p, g, b, 2, c0, P0, N0, Pn, Nn);
    if (lane == 5 && h1 && hn)
        nr += pairs_cross(p, g, b, 2, c1, P1, N1, Pn, Nn);
    if (nr)
        atomicAdd(&b.r_count, (unsigned long long)nr);
}

// Stream this CTA's histogram segment: tombstone entries touching c0/c1 (FreqMap::erase_if,
// state_opr.cc:291-294) and return the best surviving candidate.  When the segment has gone
// stale enough it is compacted in place during the same pass.
__device__ Best scan_segment(const GroupCtx &g, BlockCtx &b, uint32_t c0, uint32_t c1, bool purge, uint32_t thresh, long long *compactions) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int len = b.seg_len;
    Best best{0u, 0u, 0u};
    const bool compact = (len - b.seg_live) > (b.seg_live >> 1) + 2048 || len > g.ws.fseg_cap - (g.ws.fseg_cap >> 3);
    __syncthreads();
    if (!compact) {
        int removed = 0;
        for (int i = tid; i < len; i += nt) {
            FEnt e = g.seg[i];
            if (e.y == 0)
                continue;
            uint64_t key = ((uint64_t)e.w << 32) | e.z;
            uint32_t a = key_id0(key), c = key_id1(key);
            if (purge && (a == c0 || a == c1 || c == c0 || c == c1)) {
                g.seg[i].y = 0;
                ++removed;
                continue;
            }
            if (e.x >= thresh) {
                Best cand{e.x, e.w, e.z};
                if (best_gt(cand, best))
                    best = cand;
            }
        }
        if (removed)
            atomicSub(&b.seg_live, removed);
        __syncthreads();
        return best;
    }
    // compacting pass: tiles of nt entries, in-place (write position never passes read position)
    int out_base = 0;
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    for (int base = 0; base < len; base += nt) {
        int i = base + tid;
        FEnt e = make_uint4(0, 0, 0, 0);
        bool live = false;
        if (i < len) {
            e = g.seg[i];
            if (e.y != 0) {
                uint64_t key = ((uint64_t)e.w << 32) | e.z;
                uint32_t a = key_id0(key), c = key_id1(key);
                live = !(purge && (a == c0 || a == c1 || c == c0 || c == c1));
            }
        }
        unsigned bal = __ballot_sync(0xffffffffu, live);
        if (lane == 0)
            b.warp_sum[wid] = __popc(bal);
        __syncthreads(); // all reads of this tile done, warp sums visible
        int pre = 0, tot = 0;
        for (int w = 0; w < nw; ++w) {
            int s = b.warp_sum[w];
            if (w < wid)
                pre += s;
            tot += s;
        }
        if (live) {
            int pos = out_base + pre + __popc(bal & ((1u << lane) - 1u));
            g.seg[pos] = e;
            if (e.x >= thresh) {
                Best cand{e.x, e.w, e.z};
                if (best_gt(cand, best))
                    best = cand;
            }
        }
        out_base += tot;
        __syncthreads();
    }
    if (tid == 0) {
        b.seg_len = out_base;
        b.seg_live = out_base;
        if (compactions)
            atomicAdd((unsigned long long *)compactions, 1ull);
    }
    __syncthreads();
    return best;
}

// Block-reduce the per-thread best and publish it in this CTA's slot.
__device__ void publish_best(const GroupCtx &g, BlockCtx &b, Best best, int parity) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    best = warp_best(best);
    if (lane == 0)
        b.warp_best[wid] = best;
    __syncthreads();
    if (wid == 0) {
        Best v = lane < nw ? b.warp_best[lane] : Best{0u, 0u, 0u};
        v = warp_best(v);
        if (lane == 0) {
            uint4 s;
            s.x = v.score;
            s.y = v.khi;
            s.z = v.klo;
            s.w = (uint32_t)b.seg_live | ((uint32_t)b.status << 28);
            if (g.G > 1)
                __stcg(&g.ws.slots[parity * g.G + g.rank], s);
            else
                g.ws.slots[parity] = s;
        }
    }
}

// After a group barrier: combine all slots -> chosen pair (identical on every CTA); returns |F|.
__device__ int collect_best(const GroupCtx &g, BlockCtx &b, int parity) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    Best v{0u, 0u, 0u};
    int live = 0, st = 0;
    if (g.G > 1) {
        for (int i = tid; i < g.G; i += blockDim.x) {
            uint4 s = __ldcg(&g.ws.slots[parity * g.G + i]);
            Best c{s.x, s.y, s.z};
            if (best_gt(c, v))
                v = c;
            live += (int)(s.w & 0x0fffffffu);
            st = max(st, (int)(s.w >> 28));
        }
    }
    else if (tid == 0) {
        uint4 s = g.ws.slots[parity];
        v = Best{s.x, s.y, s.z};
        live = (int)(s.w & 0x0fffffffu);
        st = (int)(s.w >> 28);
    }
    v = warp_best(v);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        live += __shfl_xor_sync(0xffffffffu, live, off);
        st = max(st, __shfl_xor_sync(0xffffffffu, st, off));
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
            s2 = max(s2, __shfl_xor_sync(0xffffffffu, s2, off));
        }
        if (lane == 0) {
            b.chosen = w;
            b.scratch_i[0] = l;
            b.scratch_i[1] = s2;
        }
    }
    __syncthreads();
    return b.scratch_i[0];
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
struct PopResult {
    HeapEnt e;
};
__device__ __forceinline__ HeapEnt shfl_ent(const HeapEnt &v, int src) {
    HeapEnt o;
    o.lat = __shfl_sync(0xffffffffu, v.lat, src);
    o.qmin = __shfl_sync(0xffffffffu, v.qmin, src);
    o.qmax = __shfl_sync(0xffffffffu, v.qmax, src);
    o.qstep = __shfl_sync(0xffffffffu, v.qstep, src);
    o.sub = __shfl_sync(0xffffffffu, v.sub, src);
    o.la = __shfl_sync(0xffffffffu, v.la, src);
    o.id = __shfl_sync(0xffffffffu, v.id, src);
    o.shift = __shfl_sync(0xffffffffu, v.shift, src);
    return o;
}
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
        int ol = __shfl_xor_sync(0xffffffffu, wl, off);
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

__device__ void column_finish(const ProblemDesc &p, const GroupCtx &g, int o, int gid_base) {
    const int lane = threadIdx.x & 31;
    const ColEnt *list = g.ws.col_ents + (size_t)o * p.col_cap;
    const int L = g.ws.col_len[o];
    uint4 *hl = g.ws.heap + 2 * ((size_t)o * 32 + lane) * (size_t)p.heap_lane_cap;
    // digits of the rows k = lane (mod 32) go to this lane's private list
    int cnt = 0;
    for (int k = lane; k < L; k += 32) {
        ColEnt c = list[k];
        for (uint32_t m = c.P | c.N; m; m &= m - 1) {
            int sh = __ffs(m) - 1;
            QInt q;
            float lat;
            load_op(p, c.e, q, lat);
            HeapEnt e;
            e.lat = lat;
            e.sub = (int)((c.N >> sh) & 1);
            e.la = left_align(q, sh);
            e.qmin = q.min;
            e.qmax = q.max;
            e.qstep = q.step;
            e.id = (int)c.e;
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
        }
        return;
    }
    int gid = gid_base;
    while (n > 1) {
        HeapEnt e0 = heap_pop(hl, cnt);
        HeapEnt e1 = heap_pop(hl, cnt);
        // every lane holds (e0, e1): compute the merged entry redundantly, lane 0 records the op
        QInt q0{e0.qmin, e0.qmax, e0.qstep}, q1{e1.qmin, e1.qmax, e1.qstep};
        QInt q;
        float dlat, dcost;
        int4 misc;
        int rshift;
        if (e0.sub) {
            long long s = (long long)e0.shift - e1.shift;
            q = qint_add(q1, q0, s, e1.sub != 0, e0.sub != 0);
            cost_add(q1, q0, s, (1 ^ e1.sub) != 0, p.adder_size, p.carry_size, dlat, dcost);
            misc = make_int4(e1.id, e0.id, 1 ^ e1.sub, (int)s);
            rshift = e1.shift;
        }
        else {
            long long s = (long long)e1.shift - e0.shift;
            q = qint_add(q0, q1, s, e0.sub != 0, e1.sub != 0);
            cost_add(q0, q1, s, e1.sub != 0, p.adder_size, p.carry_size, dlat, dcost);
            misc = make_int4(e0.id, e1.id, e1.sub, (int)s);
            rshift = e0.shift;
        }
        float lat = fadd(fmaxf_std(e0.lat, e1.lat), dlat);
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
        HeapEnt e = heap_load(hl, 0);
        p.out_idx[o] = e.id;
        p.out_neg[o] = e.sub;
        p.out_shift[o] = base_shift + e.shift;
    }
}

// ------------------------------------------------------------------------------------------------
// solve one problem with the CTAs of one group

__device__ void solve_problem(const ProblemDesc &p, const GroupCtx &g, BlockCtx &b) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int n_in = p.n_in, n_out = p.n_out, nbits = p.nbits;
    const uint32_t thresh = method_threshold(p.method);

    if (tid == 0) {
        b.seg_len = 0;
        b.seg_live = 0;
        b.touch_n = 0;
        b.status = ST_OK;
        b.r_count = 0ull;
        b.chosen = Best{0u, 0u, 0u};
    }
    __syncthreads();

    // ---- column lists (state_opr.cc:100-112): warp per owned column, rows in ascending expr order
    for (int oc = g.rank + g.G * wid; oc < n_out; oc += g.G * nw) {
        ColEnt *list = g.ws.col_ents + (size_t)oc * p.col_cap;
        int len = 0;
        for (int i0 = 0; i0 < n_in; i0 += 32) {
            int i = i0 + lane;
            uint2 m = i < n_in ? p.masks0[(size_t)i * n_out + oc] : make_uint2(0, 0);
            bool has = (m.x | m.y) != 0;
            unsigned bal = __ballot_sync(0xffffffffu, has);
            if (has) {
                int pos = len + __popc(bal & ((1u << lane) - 1u));
                if (pos < p.col_cap) {
                    ColEnt ce{(uint32_t)i, m.x, m.y, 0u};
                    list[pos] = ce;
                }
            }
            len += __popc(bal);
        }
        if (lane == 0)
            g.ws.col_len[oc] = min(len, p.col_cap);
    }
    // ---- input ops (state_opr.cc:146-149)
    for (int i = g.rank * nt + tid; i < n_in; i += g.G * nt) {
        p.op_misc[i] = make_int4(i, -1, -1, 0);
        p.op_q[i] = make_float4(p.qint[3 * i], p.qint[3 * i + 1], p.qint[3 * i + 2], p.lat[i]);
        p.op_cost[i] = 0.0f;
    }

    Best best{0u, 0u, 0u};
    unsigned long long r0 = 0;
    if (p.method != M_DUMMY) {
        // ---- initial pair histogram (state_opr.cc:115-144, types.hh:73-100): warp per (a<=b) pair block,
        //      lanes over relative shifts, sign planes streamed over the output columns
        const long long n_pairs = (long long)n_in * (n_in + 1) / 2;
        const int n_sh = 2 * nbits - 1;
        for (long long pi = (long long)g.rank * nw + wid; pi < n_pairs; pi += (long long)g.G * nw) {
            // decode pi -> (a, c) with a <= c, row-major over the upper triangle
            long long a = (long long)(((2.0 * n_in + 1.0) - sqrt((2.0 * n_in + 1.0) * (2.0 * n_in + 1.0) - 8.0 * (double)pi)) * 0.5);
            while (a > 0 && a * (2LL * n_in - a + 1) / 2 > pi)
                --a;
            while ((a + 1) * (2LL * n_in - (a + 1) + 1) / 2 <= pi)
                ++a;
            long long c = a + (pi - a * (2LL * n_in - a + 1) / 2);
            const uint2 *ra = p.masks0 + (size_t)a * n_out;
            const uint2 *rc = p.masks0 + (size_t)c * n_out;
            for (int s0 = 0; s0 < n_sh; s0 += 32) {
                int si = s0 + lane;
                int s = si - (nbits - 1);
                bool active = si < n_sh && !(a == c && s >= 0);
                uint32_t same = 0, diff = 0;
                if (active) {
                    if (s >= 0) {
                        for (int o = 0; o < n_out; ++o) {
                            uint2 x = ra[o], y = rc[o];
                            same += __popc(x.x & (y.x >> s)) + __popc(x.y & (y.y >> s));
                            diff += __popc(x.x & (y.y >> s)) + __popc(x.y & (y.x >> s));
                        }
                    }
                    else {
                        int d = -s;
                        for (int o = 0; o < n_out; ++o) {
                            uint2 x = ra[o], y = rc[o];
                            same += __popc((x.x >> d) & y.x) + __popc((x.y >> d) & y.y);
                            diff += __popc((x.x >> d) & y.y) + __popc((x.y >> d) & y.x);
                        }
                    }
                    r0 += same + diff;
                    if (same >= 2)
                        emit_entry(p, g, b, (uint32_t)a, (uint32_t)c, s, 0, same, thresh, best);
                    if (diff >= 2)
                        emit_entry(p, g, b, (uint32_t)a, (uint32_t)c, s, 1, diff, thresh, best);
                }
            }
        }
    }
    if (r0)
        atomicAdd(&b.r_count, r0);
    __syncthreads();
    const unsigned long long r0_cta = b.r_count;
    int parity = 0;
    publish_best(g, b, best, parity);
    group_sync(g, b);
    int f_live = collect_best(g, b, parity);
    parity ^= 1;
    const int f0 = f_live;
    int f_max = f_live;

    // ---- greedy loop (cmvm_core.cc:36-70)
    int t = 0;
    unsigned long long sum_f = 0;
    int status = b.scratch_i[1];
    while (status == ST_OK) {
        Best ch = b.chosen;
        if (ch.score == 0u || p.method == M_DUMMY)
            break;
        if (n_in + t >= p.e_cap) {
            status = ST_EXPR_OVERFLOW;
            break;
        }
        const uint64_t key = ((uint64_t)ch.khi << 32) | ch.klo;
        const uint32_t c0 = key_id0(key), c1 = key_id1(key);
        const int shift = key_shift(key), sub = key_sub(key);
        const uint32_t newid = (uint32_t)(n_in + t);
        sum_f += (unsigned long long)f_live;
        f_max = max(f_max, f_live);
        if (g.rank == 0 && tid == 0) {
            // pair_to_op (state_opr.cc:211-225)
            QInt q0, q1;
            float l0, l1;
            load_op(p, c0, q0, l0);
            load_op(p, c1, q1, l1);
            float dlat, cost;
            cost_add(q0, q1, shift, sub != 0, p.adder_size, p.carry_size, dlat, cost);
            QInt q = qint_add(q0, q1, shift, false, sub != 0);
            float lat = fadd(fmaxf_std(l0, l1), dlat);
            p.op_misc[newid] = make_int4((int)c0, (int)c1, sub, shift);
            p.op_q[newid] = make_float4(q.min, q.max, q.step, lat);
            p.op_cost[newid] = cost;
            if (p.trace && t < p.trace_cap) {
                int *tr = p.trace + 5 * (size_t)t;
                tr[0] = (int)c0;
                tr[1] = (int)c1;
                tr[2] = shift;
                tr[3] = sub;
                tr[4] = f_live;
            }
        }
        // A. substitute + recount in the owned columns
        for (int oc = g.rank + g.G * wid; oc < n_out; oc += g.G * nw)
            column_step(p, g, b, oc, c0, c1, shift, sub, newid);
        group_arrive(g, b);
        // B. purge + scan of the old entries overlaps the other CTAs' column work
        best = scan_segment(g, b, c0, c1, true, thresh, g.rank == 0 ? &p.result_meta[META_COMPACTIONS] : nullptr);
        group_wait(g, b);
        // C. harvest the counters this CTA touched first
        const int n_touch = min(b.touch_n, g.ws.touch_cap);
        for (int i = tid; i < n_touch; i += nt) {
            uint32_t idx = g.touch[i];
            uint32_t cnt = __ldcg(&g.ws.slab[idx]);
            g.ws.slab[idx] = 0u;
            if (cnt >= 2) {
                int sb = (int)(idx & 1u);
                int si = (int)((idx >> 1) & ((1u << (p.log_s - 1)) - 1u));
                uint32_t r = idx >> p.log_s;
                int slot = r >= 2u * (uint32_t)p.e_cap ? 2 : (r >= (uint32_t)p.e_cap ? 1 : 0);
                uint32_t x = r - (uint32_t)slot * (uint32_t)p.e_cap;
                uint32_t m = slot == 0 ? c0 : (slot == 1 ? c1 : newid);
                emit_entry(p, g, b, min(m, x), max(m, x), si - (nbits - 1), sb, cnt, thresh, best);
            }
        }
        __syncthreads();
        if (tid == 0)
            b.touch_n = 0;
        publish_best(g, b, best, parity);
        group_sync(g, b);
        f_live = collect_best(g, b, parity);
        parity ^= 1;
        ++t;
        if (b.scratch_i[1] != ST_OK) { // some CTA overflowed a buffer: every CTA sees it in the slots and stops
            status = b.scratch_i[1];
            break;
        }
    }

    // ---- to_solution
    for (int oc = g.rank + g.G * wid; oc < n_out; oc += g.G * nw) {
        const ColEnt *list = g.ws.col_ents + (size_t)oc * p.col_cap;
        const int L = g.ws.col_len[oc];
        int k = 0;
        for (int i = lane; i < L; i += 32)
            k += __popc(list[i].P) + __popc(list[i].N);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            k += __shfl_xor_sync(0xffffffffu, k, off);
        if (lane == 0) {
            if (g.G > 1)
                __stcg(&g.ws.col_k[oc], k);
            else
                g.ws.col_k[oc] = k;
        }
    }
    group_sync(g, b);
    for (int oc = g.rank + g.G * wid; oc < n_out; oc += g.G * nw) {
        int before = 0;
        for (int i = lane; i < oc; i += 32) {
            int k = __ldcg(&g.ws.col_k[i]);
            before += k > 1 ? k - 1 : 0;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            before += __shfl_xor_sync(0xffffffffu, before, off);
        column_finish(p, g, oc, n_in + t + before);
    }
    // ---- bookkeeping
    __syncthreads();
    if (tid == 0) {
        atomicAdd((unsigned long long *)&p.result_meta[META_SUM_R], b.r_count - r0_cta);
        atomicAdd((unsigned long long *)&p.result_meta[META_R0], r0_cta);
        if (b.status != ST_OK)
            atomicMax((int *)&p.result_meta[META_STATUS], b.status);
    }
    if (g.rank == 0 && tid == 0) {
        long long tree = 0, dfin = 0;
        for (int o = 0; o < n_out; ++o) {
            int k = __ldcg(&g.ws.col_k[o]);
            tree += k > 1 ? k - 1 : 0;
            dfin += k;
        }
        long long n_ops = (long long)n_in + t + tree;
        p.result_meta[META_N_OPS] = n_ops;
        p.result_meta[META_T] = t;
        p.result_meta[META_SUM_F] = (long long)sum_f;
        p.result_meta[META_F0] = f0;
        p.result_meta[META_D_FINAL] = dfin;
        p.result_meta[META_F_MAX] = f_max;
        if (status != ST_OK)
            atomicMax((int *)&p.result_meta[META_STATUS], status);
        if (n_ops > p.ops_cap)
            atomicMax((int *)&p.result_meta[META_STATUS], (int)ST_OPS_OVERFLOW);
    }
    group_sync(g, b); // workspace may be reused by the next problem of this group
}

// grid = n_groups * G CTAs; group i solves problems i, i + n_groups, ...
__global__ void __launch_bounds__(512, 1) cmvm_solve_kernel(const ProblemDesc *probs, int n_probs, const GroupWs *wss, int G) {
    __shared__ BlockCtx b;
    GroupCtx g;
    g.G = G;
    g.rank = blockIdx.x % G;
    const int group = blockIdx.x / G, n_groups = gridDim.x / G;
    g.ws = wss[group];
    g.seg = g.ws.fseg + (size_t)g.rank * g.ws.fseg_cap;
    g.touch = g.ws.touch + (size_t)g.rank * g.ws.touch_cap;
    if (threadIdx.x == 0)
        b.bar_target = 0u; // the host zeroes the arrive counter before every launch
    __syncthreads();
    for (int pi = group; pi < n_probs; pi += n_groups)
        solve_problem(probs[pi], g, b);
}

} // namespace da
