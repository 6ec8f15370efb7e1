// solve_columns.cuh -- substitution of the chosen pair in the owned output columns and the recount of digit pairs
// (update_expr state_opr.cc:227-283, the per-column share of update_stats state_opr.cc:307-340).
#pragma once
#include "solve_common.cuh"

namespace da {

// ------------------------------------------------------------------------------------------------
// solve: pieces

__device__ __forceinline__ uint32_t slab_index(const ProblemDesc &p, int slot, uint32_t x, int shift, int sub) {
    return (((uint32_t)slot * (uint32_t)p.e_cap + x) << p.log_s) + (uint32_t)(((shift + p.nbits - 1) << 1) | sub);
}

// first toucher of a counter records it: harvesting is O(distinct pairs) and leaves the slab zero
__device__ __forceinline__ void touch_push(const Ctx &cx, uint32_t idx) {
    const int t = smem_add(&cx.b->touch_n, 1);
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
        const int a = smem_add(&cx.b->n_act, 1);
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
    // items ascend with the loop, so each thread walks the list of touched columns once (cursor = column a, first item
    // `start` of that column, its length `span`)
    int a = 0, start = 0, span = *col_ref(cx, p, cx.act[0].slot, cx.act[0].o).len;
    for (int item = tid; item < total_pad; item += nt) {
        PairSrc s0, s1, s2;
        s0.n = s1.n = s2.n = 0;
        s0.dhi = s1.dhi = s2.dhi = 1;
        if (item < total) {
            while (item >= start + span) {
                start += span;
                ++a;
                span = *col_ref(cx, p, cx.act[a].slot, cx.act[a].o).len;
            }
            const int k = item - start;
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
        smem_add(&b.r_step, nr);
}


} // namespace da
