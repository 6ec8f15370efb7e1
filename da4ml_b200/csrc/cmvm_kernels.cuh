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
#include "cmvm_prep.cuh"
#include "solve_columns.cuh"
#include "solve_finish.cuh"
#include "solve_histogram.cuh"

namespace da {

// ------------------------------------------------------------------------------------------------
// solve one problem with the CTAs of one group

__device__ void solve_problem(const ProblemDesc &p, const Ctx &cx) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int n_in = p.n_in, n_out = p.n_out, nbits = p.nbits, G = cx.cfg.G;
    const uint32_t thresh = method_threshold(p.method);
    BlockCtx &b = *cx.b;

    if (tid == 0) {
        b.seg_len = 0;
        b.cap0 = cx.ws.fseg_cap;
        b.hot_n = 0;
        b.hot_cap = 0;
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
    // every CTA's share of the stamps is zero before anybody judges an entry by them (the initial refresh below runs
    // before the first exchange; found by the race check of the CPU kernel simulation)
    group_sync(cx);

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
    if (p.method != M_DUMMY)
        r0 = initial_histogram(p, cx, thresh, best);
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
    refresh_chunks(cx, team_all(), b.seg_len, 0u, 0u, false, true, thresh);
    publish_best(cx, Best{0u, 0u, 0u}); // the exchange is also the barrier that orders mod_step zeroing / list building
    int f_live = collect_best(cx);
    const int f0 = f_live;
    int f_max = f_live;

    // ---- greedy loop (cmvm_core.cc:36-70)
    int t = 0;
    unsigned long long sum_f = 0;
    int status = b.scratch_i[1];
    const long long t_start = clock64();
    while (status == ST_OK) {
        const Best ch = b.chosen;
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
            st_racy(&cx.ws.mod_step[c0], stamp);
            st_racy(&cx.ws.mod_step[c1], stamp);
            st_racy(&cx.ws.mod_step[newid], stamp);
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
            refresh_chunks(cx, team_all(), b.seg_len, c0, c1, true, cx.cfg.accounting != 0, thresh);
        DA_LAP(2)
        // (read before anybody appends in this step: threads that are already harvesting must not change what the
        // slower ones decide below -- found by the CPU kernel simulation with a deliberately small segment)
        const int seg_before = b.seg_len;
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
        if (seg_before + n_mine > cx.ws.fseg_cap) // (uniform over the CTA)
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
        if (cx.rank == 0 && tid == 0 && t >= 250 && t % 250 == 0 && ((t / 250) & (t / 250 - 1)) == 0 && t / 250 <= 64) {
            long long *ms = &p.result_meta[META_MILESTONES + 9 * (31 - __clz(t / 250))]; // (diagnostic: where the time of a stage goes)
            for (int k = 0; k < 8; ++k)
                ms[k] = b.phase[k];
            ms[8] = clock64() - t_start;
        }
        if (b.scratch_i[1] != ST_OK) { // some CTA overflowed a buffer: every CTA sees it in the exchange and stops
            status = b.scratch_i[1];
            break;
        }
    }

    // ---- to_solution
    finish_columns(p, cx, t);
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
            atomicMax((long long *)&p.result_meta[META_PHASEMAX + k], b.phase[k]); // the slowest CTA's total per phase
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
    DA_DYN_SHARED(smem);
    DA_SHARED_VAR(BlockCtx, bctx);
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

} // namespace da

