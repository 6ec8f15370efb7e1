// cmvm_kernel_em.cuh -- persistent solve kernel in the expression-major formulation of solve_rows.cuh (EXPERIMENTAL:
// exercised by the CPU kernel simulation in tests/, not yet launched by the host driver).  Same results as
// cmvm_solve_kernel; per greedy step ONE group exchange (the argmax) and no cross-CTA counters.
#pragma once
#include "cmvm_kernels.cuh"
#include "solve_rows.cuh"

namespace da {

__device__ void solve_problem_em(const ProblemDesc &p, const Ctx &cx, const EmCtx &ex) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int n_in = p.n_in, n_out = p.n_out, nbits = p.nbits, G = cx.cfg.G;
    const uint32_t thresh = method_threshold(p.method);
    BlockCtx &b = *cx.b;
    EmBlock &eb = *ex.eb;

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
        for (int k = 0; k < 8; ++k) {
            b.phase[k] = 0;
            b.peak[k] = 0;
            b.nslow[k] = 0;
        }
        b.poll_iters = 0;
        eb.pool_used = 0;
        eb.tile_n = 0;
    }
    for (int c = tid; c < cx.cfg.nchunk_cap; c += nt) {
        cx.cb_score[c] = 0u;
        cx.cb_khi[c] = 0u;
        cx.cb_klo[c] = 0u;
        cx.cb_dirty[c] = 0;
    }
    for (int i = cx.rank * nt + tid; i < p.e_cap; i += G * nt)
        cx.ws.mod_step[i] = 0u;
    {
        unsigned char *ver = ex.ws.ver + (size_t)cx.rank * ex.ws.e_cap; // this CTA's private replica
        for (int i = tid; i < p.e_cap; i += nt)
            ver[i] = 0;
    }
    group_sync(cx); // every CTA's share of the stamps is zero before anybody judges an entry by them
    em_init_cells(p, cx, ex);
    // ---- input ops (state_opr.cc:146-149)
    for (int i = cx.rank * nt + tid; i < n_in; i += G * nt) {
        p.op_misc[i] = make_int4(i, -1, -1, 0);
        p.op_q[i] = make_float4(p.qint[3 * i], p.qint[3 * i + 1], p.qint[3 * i + 2], p.lat[i]);
        p.op_cost[i] = 0.0f;
    }

    Best best{0u, 0u, 0u};
    unsigned long long r0 = 0;
    if (p.method != M_DUMMY) {
        // ---- initial pair histogram (state_opr.cc:115-144): as in cmvm_solve_kernel
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
    refresh_chunks(cx, 0u, 0u, false, true, thresh);
    publish_best(cx, Best{0u, 0u, 0u}); // the exchange also orders the cells / input ops written above
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
        if (tid == nt - 1) {
            // pair_to_op (state_opr.cc:211-225): every CTA needs the new record for the entries it emits in this step;
            // CTA 0 also publishes it (with the rewrite stamps) for the later steps
            QInt q0, q1;
            float l0, l1;
            load_op(p, c0, q0, l0);
            load_op(p, c1, q1, l1);
            float dlat, cost;
            cost_add(q0, q1, shift, sub != 0, p.adder_size, p.carry_size, dlat, cost);
            const QInt q = qint_add(q0, q1, shift, false, sub != 0);
            const float lat = fadd(fmaxf_std(l0, l1), dlat);
            int r = 0;
            eb.mid[r] = c0, eb.mq[r] = q0, eb.ml[r] = l0, ++r;
            if (c1 != c0)
                eb.mid[r] = c1, eb.mq[r] = q1, eb.ml[r] = l1, ++r;
            eb.mid[r] = newid, eb.mq[r] = q, eb.ml[r] = lat, ++r;
            eb.n_mods = r;
            if (cx.rank == 0) {
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
        }
        if (tid == 0) {
            b.t_last = clock64();
            b.live_old = 0;
            b.n_new = 0;
            b.r_count += (unsigned long long)b.r_step;
            b.r_step = 0;
            b.rescanned += (unsigned long long)b.rescan_step;
            b.rescan_step = 0;
        }
        // A. substitution in every column (redundantly on every CTA), B. the owners update their cells
        em_substitute(p, cx, ex, c0, c1, shift, sub);
        em_update_owned(p, cx, ex, c0, c1, newid);
        DA_LAP(0)
        // C. argmax caches first (entries touching c0 / c1 die), then the recount appends this step's entries
        if (b.scratch_i[3])
            compact_segment(cx, c0, c1, true, thresh, cx.rank == 0 ? &p.result_meta[META_COMPACTIONS] : nullptr);
        else
            refresh_chunks(cx, c0, c1, true, cx.cfg.accounting != 0, thresh);
        DA_LAP(2)
        best = Best{0u, 0u, 0u};
        em_recount(p, cx, ex, newid, stamp, thresh, best);
        __syncthreads();
        DA_LAP(1)
        publish_best(cx, best);
        DA_LAP(5)
        f_live = collect_best(cx);
        DA_LAP(7)
        ++t;
        if (b.scratch_i[1] != ST_OK) {
            status = b.scratch_i[1];
            break;
        }
    }

    // ---- to_solution: column lists from the cells, then exactly as in cmvm_solve_kernel
    em_scatter_columns(p, cx, ex, (uint32_t)(n_in + t));
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
    group_sync(cx);
    if (cx.rank == 0 && wid == 0) {
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
    __syncthreads();
    if (tid == 0) {
        b.r_count += (unsigned long long)b.r_step;
        b.rescanned += (unsigned long long)b.rescan_step;
        atomicAdd((unsigned long long *)&p.result_meta[META_SUM_R], b.r_count - r0_cta);
        atomicAdd((unsigned long long *)&p.result_meta[META_R0], r0_cta);
        atomicAdd((unsigned long long *)&p.result_meta[META_RESCANNED], b.rescanned);
        atomicMax((long long *)&p.result_meta[META_LIST_MAX], (long long)eb.pool_used);
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
        if (status != ST_OK)
            atomicMax((int *)&p.result_meta[META_STATUS], status);
        if (n_ops > p.ops_cap)
            atomicMax((int *)&p.result_meta[META_STATUS], (int)ST_OPS_OVERFLOW);
    }
    group_sync(cx);
}

// shared memory of one CTA: chunk caches (as cmvm_solve_kernel) + three dense rows, their bitmaps, the tile list
__host__ __device__ inline size_t em_smem_bytes(int nchunk_cap, int n_out_max, int cta_threads) {
    const int words = (n_out_max + 31) / 32;
    return (size_t)nchunk_cap * 17 + 64 + sizeof(uint2) * 3 * (size_t)n_out_max + sizeof(uint32_t) * (5 * (size_t)words + 3 * (size_t)cta_threads) + sizeof(float4) * (size_t)cta_threads + 64;
}

__device__ __forceinline__ void solve_em_kernel_body(const ProblemDesc *probs, int n_probs, const GroupWs *wss, const EmWs *ews, const LaunchCfg &cfg, int n_out_max) {
    DA_DYN_SHARED(smem);
    DA_SHARED_VAR(BlockCtx, bctx);
    DA_SHARED_VAR(EmBlock, eblk);
    Ctx cx;
    cx.cfg = cfg;
    cx.rank = blockIdx.x % cfg.G;
    const int group = blockIdx.x / cfg.G, n_groups = gridDim.x / cfg.G;
    cx.ws = wss[group];
    cx.seg = cx.ws.fseg + (size_t)cx.rank * cx.ws.fseg_cap;
    cx.touch_g = nullptr;
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
    cx.col_len_s = nullptr;
    cx.act = nullptr;
    cx.lists_s = nullptr;
    EmCtx ex;
    ex.ws = ews[group];
    ex.eb = &eblk;
    const int words = (n_out_max + 31) / 32;
    sp = (unsigned char *)(((uintptr_t)sp + 15) & ~(uintptr_t)15);
    for (int r = 0; r < 3; ++r) {
        ex.D[r] = (uint2 *)sp;
        sp += sizeof(uint2) * n_out_max;
    }
    for (int r = 0; r < 3; ++r) {
        ex.B[r] = (uint32_t *)sp;
        sp += sizeof(uint32_t) * words;
    }
    ex.A = (uint32_t *)sp;
    sp += sizeof(uint32_t) * words;
    ex.pre = (uint32_t *)sp;
    sp += sizeof(uint32_t) * words;
    sp = (unsigned char *)(((uintptr_t)sp + 15) & ~(uintptr_t)15);
    ex.tile_q = (float4 *)sp;
    sp += sizeof(float4) * blockDim.x;
    ex.tile = (uint32_t *)sp;
    sp += sizeof(uint32_t) * blockDim.x;
    ex.tile_off = (uint32_t *)sp;
    sp += sizeof(uint32_t) * blockDim.x;
    ex.tile_cnt = (uint32_t *)sp;
    sp += sizeof(uint32_t) * blockDim.x;
    cx.cb_dirty = sp;
    if (threadIdx.x == 0) {
        bctx.bar_target = 0u;
        bctx.epoch = 0u;
    }
    __syncthreads();
    for (int pi = group; pi < n_probs; pi += n_groups)
        solve_problem_em(probs[pi], cx, ex);
}
__global__ void __launch_bounds__(512, 1) cmvm_solve_em_kernel(const ProblemDesc *probs, int n_probs, const GroupWs *wss, const EmWs *ews, LaunchCfg cfg, int n_out_max) {
    solve_em_kernel_body(probs, n_probs, wss, ews, cfg, n_out_max);
}

} // namespace da
