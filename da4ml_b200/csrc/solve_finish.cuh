// solve_finish.cuh -- adder-tree finisher, one warp per output column with lane-private heaps (to_solution,
// cmvm_core.cc:75-225).
#pragma once
#include "solve_common.cuh"

namespace da {

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

// the rows of column o: expression id and sign planes, [3][col_cap] in global memory (written by own_scatter_columns)
struct ColList {
    const uint32_t *e, *P, *N;
    int len;
};
__device__ __forceinline__ ColList col_list(const Ctx &cx, const ProblemDesc &p, int o) {
    const uint32_t *base = cx.ws.col_u32 + (size_t)o * 3 * p.col_cap;
    return ColList{base, base + p.col_cap, base + 2 * p.col_cap, cx.ws.col_len[o]};
}
__device__ __noinline__ void column_finish(const ProblemDesc &p, const Ctx &cx, int o, int gid_base) {
    const int lane = threadIdx.x & 31;
    const ColList L = col_list(cx, p, o);
    const int len = L.len;
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


// to_solution for every column (cmvm_core.cc:89-225): digits per column -> op ids of the adder trees (sequential across
// columns, cmvm_core.cc:101,203) -> one warp per column; then the float cost of the stage, summed in op order
// (api.cc:222-227).  Group-wide; `t` = greedy steps done (the trees' ops follow the n_in + t records of the loop).
__device__ void finish_columns(const ProblemDesc &p, const Ctx &cx, int t) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n_in = p.n_in, n_out = p.n_out, G = cx.cfg.G;
    for (int slot = wid; slot < cx.cfg.cpc; slot += nw) {
        const int oc = cx.rank + G * slot;
        if (oc >= n_out)
            break;
        const ColList L = col_list(cx, p, oc);
        const int len = L.len;
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
        column_finish(p, cx, oc, n_in + t + before);
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
}

} // namespace da
