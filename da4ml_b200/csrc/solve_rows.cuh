// solve_rows.cuh -- expression-major form of the greedy step (EXPERIMENTAL: exercised by the CPU kernel simulation in
// tests/, not yet wired into the host driver or measured on a GPU).
//
// Expression e is owned by CTA e mod G.  The owner keeps e's live digits as a list of cells (column, P, N) plus a column
// bitmap.  A step then needs no cross-CTA counters at all:
//   * every CTA rebuilds the dense rows of c0 and c1 in shared memory from their cells and computes the substitution
//     for all columns itself (update_expr, state_opr.cc:227-283) -> dense rows c0', c1', new;
//   * the owners of c0, c1 and of the new expression rewrite / create their cells;
//   * every CTA recounts, for each expression x it owns, the pairs (m, x), m in {c0, c1, new} (update_stats,
//     state_opr.cc:307-340) with popcounts over x's cells -- lanes of a warp are the relative shifts -- and appends
//     the entries with count >= 2 to its own histogram segment.
// One group exchange per step (the argmax) instead of two, no counter slab, no touched lists, no harvest.
#pragma once
#include "solve_common.cuh"

namespace da {

struct EmWs {
    // cells of the expressions a CTA owns, bump-allocated from its slice [rank * pool_cap, (rank + 1) * pool_cap):
    // column + two versions of the sign planes.  Other CTAs read the cells of c0 / c1 at the start of a step while the
    // owner rewrites them later in the same step, so a rewrite goes to the other version; which version is current is
    // the parity of the number of rewrites of the expression, which every CTA tracks privately (ver).
    uint32_t *cell_col; // [G * pool_cap]
    uint2 *cell_pl[2];  // [G * pool_cap] each
    uint32_t *cell_off; // [e_cap] first cell of an expression (absolute index)
    uint32_t *cell_cnt; // [e_cap] cells allocated to it (dead cells keep empty planes)
    uint32_t *rowbits;  // [e_cap][words] columns in which the expression has digits (read by the owner only)
    float4 *own_q;      // [e_cap] owner-major copy of the op record (qmin, qmax, qstep, latency), read by the owner's scan
    unsigned char *ver; // [G][e_cap] per-CTA replica: rewrites of each expression so far
    int pool_cap;       // cells per CTA
    int words;          // ceil(n_out_max / 32)
    int e_cap;
    int per;            // ceil(e_cap / G): cell_off / cell_cnt / rowbits are stored owner-major (see em_slot)
};
// Slot of expression e in the per-expression tables: owner-major, [e mod G][e / G], so that the expressions one CTA owns
// are contiguous (its scan over them is coalesced) while every CTA can still address any expression.
__device__ __forceinline__ size_t em_slot(const EmWs &ws, uint32_t e, int G) { return (size_t)(e % (uint32_t)G) * ws.per + e / (uint32_t)G; }

// per-CTA shared state of the expression-major step
struct EmBlock {
    int pool_used;  // cells handed out of this CTA's pool slice
    int tile_n;     // active expressions of the current tile
    int n_mods;     // rewritten rows of this step: 2 (self pair) or 3
    uint32_t mid[3]; // their ids, ascending: c0, (c1,) new
    QInt mq[3];      // their op records
    float ml[3];
    int new_cells;  // nonzero columns of the new row
};
struct EmCtx {
    EmWs ws;
    EmBlock *eb;
    uint2 *D[3];      // dense rows of the rewritten expressions after the substitution, [n_out] each (D[1] = D[0] for a self pair)
    uint32_t *B[3];   // their column bitmaps, [words] each
    uint32_t *A;      // union of the three
    uint32_t *pre;    // [words] exclusive prefix of popc(B[new])
    // active expressions of the current tile of blockDim.x owned expressions, with what their warp needs (fetched by the
    // scanning threads, all loads in flight together): id, first cell, cell count | (rows shared with m_r) << 28, op record
    uint32_t *tile, *tile_off, *tile_cnt;
    float4 *tile_q;
};

// update_expr for one column on sign planes (state_opr.cc:227-283); planes are updated in place, the new row returned
__device__ __forceinline__ void substitute_planes(uint32_t &P0, uint32_t &N0, uint32_t &P1, uint32_t &N1, bool self, int shift, int sub, int nbits, uint32_t &Pn, uint32_t &Nn) {
    Pn = 0u, Nn = 0u;
    if (!self) {
        const bool flip = shift < 0;
        const int rel = flip ? -shift : shift;
        const uint32_t AP = flip ? P1 : P0, AN = flip ? N1 : N0; // expr0 after the reference's swap
        const uint32_t BP = flip ? P0 : P1, BN = flip ? N0 : N1;
        const uint32_t M = sub ? ((AP & (BN >> rel)) | (AN & (BP >> rel))) : ((AP & (BP >> rel)) | (AN & (BN >> rel)));
        const uint32_t MB = M << rel;
        if (!flip) { // the new digit takes position and sign of id0's digit
            Pn = AP & M, Nn = AN & M;
            P0 = AP & ~M, N0 = AN & ~M, P1 = BP & ~MB, N1 = BN & ~MB;
        }
        else {
            Pn = BP & MB, Nn = BN & MB;
            P1 = AP & ~M, N1 = AN & ~M, P0 = BP & ~MB, N0 = BN & ~MB;
        }
    }
    else { // self pair (always shift < 0): order-dependent greedy matching with tombstones
        const int rel = -shift;
        const uint32_t live = P0 | N0;
        uint32_t tomb = 0u;
        for (uint32_t m = live; m; m &= m - 1) {
            const int pl = __ffs(m) - 1;
            if ((tomb >> pl) & 1u)
                continue;
            const int q = pl + rel;
            if (q >= nbits || q >= 32)
                continue;
            if (!((live >> q) & 1u) || ((tomb >> q) & 1u))
                continue;
            if ((int)(((N0 >> pl) ^ (N0 >> q)) & 1u) != sub)
                continue;
            if ((N0 >> q) & 1u)
                Nn |= 1u << q;
            else
                Pn |= 1u << q;
            tomb |= (1u << pl) | (1u << q);
        }
        P0 &= ~tomb, N0 &= ~tomb;
        P1 = P0, N1 = N0;
    }
}

// cells of the inputs this CTA owns (state_opr.cc:100-112), one warp per input
__device__ void em_init_cells(const ProblemDesc &p, const Ctx &cx, const EmCtx &ex) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5, G = cx.cfg.G;
    const size_t pool = (size_t)cx.rank * ex.ws.pool_cap;
    for (int i = cx.rank + G * wid; i < p.n_in; i += G * nw) {
        const uint2 *row = p.masks0 + (size_t)i * p.n_out;
        int cnt = 0;
        for (int o0 = 0; o0 < p.n_out; o0 += 32) {
            const int o = o0 + lane;
            const uint2 m = o < p.n_out ? row[o] : make_uint2(0u, 0u);
            const unsigned bal = __ballot_sync(0xffffffffu, (m.x | m.y) != 0u);
            cnt += __popc(bal);
            if (lane == 0)
                ex.ws.rowbits[em_slot(ex.ws, (uint32_t)i, G) * ex.ws.words + (o0 >> 5)] = bal;
        }
        int off = 0;
        if (lane == 0) {
            off = smem_add(&ex.eb->pool_used, cnt);
            if (off + cnt > ex.ws.pool_cap) {
                cx.b->status = ST_LIST_OVERFLOW;
                cnt = 0;
            }
            ex.ws.cell_off[em_slot(ex.ws, (uint32_t)i, G)] = (uint32_t)((size_t)cx.rank * ex.ws.pool_cap + off);
            ex.ws.cell_cnt[em_slot(ex.ws, (uint32_t)i, G)] = (uint32_t)cnt;
            ex.ws.own_q[em_slot(ex.ws, (uint32_t)i, G)] = make_float4(p.qint[3 * i], p.qint[3 * i + 1], p.qint[3 * i + 2], p.lat[i]);
        }
        off = __shfl_sync(0xffffffffu, off, 0);
        cnt = __shfl_sync(0xffffffffu, cnt, 0);
        int k = 0;
        for (int o0 = 0; o0 < p.n_out; o0 += 32) {
            const int o = o0 + lane;
            const uint2 m = o < p.n_out ? row[o] : make_uint2(0u, 0u);
            const bool on = (m.x | m.y) != 0u && cnt > 0;
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            if (on) {
                const size_t ci = pool + off + k + __popc(bal & ((1u << lane) - 1u));
                ex.ws.cell_col[ci] = (uint32_t)o;
                ex.ws.cell_pl[0][ci] = m; // version 0 (ver starts at zero)
            }
            k += __popc(bal);
        }
    }
}

// A. dense rows of c0 / c1 from their cells, the substitution in every column, bitmaps of the three new rows
__device__ void em_substitute(const ProblemDesc &p, const Ctx &cx, const EmCtx &ex, uint32_t c0, uint32_t c1, int shift, int sub) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const bool self = c0 == c1;
    uint2 *D0 = ex.D[0], *D1 = self ? ex.D[0] : ex.D[1], *Dn = ex.D[self ? 1 : 2];
    for (int o = tid; o < p.n_out; o += nt) {
        ex.D[0][o] = make_uint2(0u, 0u);
        ex.D[1][o] = make_uint2(0u, 0u);
        ex.D[2][o] = make_uint2(0u, 0u);
    }
    __syncthreads();
    {
        const unsigned char *ver = ex.ws.ver + (size_t)cx.rank * ex.ws.e_cap;
        const uint32_t off0 = __ldcg(&ex.ws.cell_off[em_slot(ex.ws, c0, cx.cfg.G)]), cnt0 = __ldcg(&ex.ws.cell_cnt[em_slot(ex.ws, c0, cx.cfg.G)]);
        const uint2 *pl0 = ex.ws.cell_pl[ver[c0] & 1];
        for (uint32_t i = tid; i < cnt0; i += nt)
            D0[__ldcg(&ex.ws.cell_col[off0 + i])] = __ldcg(&pl0[off0 + i]);
        if (!self) {
            const uint32_t off1 = __ldcg(&ex.ws.cell_off[em_slot(ex.ws, c1, cx.cfg.G)]), cnt1 = __ldcg(&ex.ws.cell_cnt[em_slot(ex.ws, c1, cx.cfg.G)]);
            const uint2 *pl1 = ex.ws.cell_pl[ver[c1] & 1];
            for (uint32_t i = tid; i < cnt1; i += nt)
                D1[__ldcg(&ex.ws.cell_col[off1 + i])] = __ldcg(&pl1[off1 + i]);
        }
    }
    __syncthreads();
    const int n_pad = (p.n_out + 31) & ~31;
    for (int o0 = wid * 32; o0 < n_pad; o0 += nw * 32) {
        const int o = o0 + lane;
        uint32_t P0 = 0u, N0 = 0u, P1 = 0u, N1 = 0u, Pn = 0u, Nn = 0u;
        if (o < p.n_out) {
            const uint2 a = D0[o], c = D1[o];
            P0 = a.x, N0 = a.y, P1 = c.x, N1 = c.y;
            if ((P0 | N0) != 0u && (P1 | N1) != 0u)
                substitute_planes(P0, N0, P1, N1, self, shift, sub, p.nbits, Pn, Nn);
            D0[o] = make_uint2(P0, N0);
            if (!self)
                D1[o] = make_uint2(P1, N1);
            Dn[o] = make_uint2(Pn, Nn);
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, (P0 | N0) != 0u);
        const unsigned b1 = __ballot_sync(0xffffffffu, (P1 | N1) != 0u);
        const unsigned bn = __ballot_sync(0xffffffffu, (Pn | Nn) != 0u);
        if (lane == 0) {
            const int w = o0 >> 5;
            ex.B[0][w] = b0;
            if (!self)
                ex.B[1][w] = b1;
            ex.B[self ? 1 : 2][w] = bn;
            ex.A[w] = b0 | b1 | bn;
        }
    }
    __syncthreads();
    if (tid == 0) { // exclusive prefix over the new row's bitmap words (cell positions of the new expression)
        const uint32_t *Bn = ex.B[self ? 1 : 2];
        int acc = 0;
        for (int w = 0; w < (p.n_out + 31) / 32; ++w) {
            ex.pre[w] = (uint32_t)acc;
            acc += __popc(Bn[w]);
        }
        ex.eb->new_cells = acc;
        // this CTA's view of the versions: c0 and c1 have now been rewritten once more (everybody read the old ones above)
        unsigned char *ver = ex.ws.ver + (size_t)cx.rank * ex.ws.e_cap;
        ver[c0] += 1;
        if (!self)
            ver[c1] += 1;
    }
    __syncthreads();
}

// B. the owners write the rewritten rows back to their cells and create the cells of the new expression
__device__ void em_update_owned(const ProblemDesc &p, const Ctx &cx, const EmCtx &ex, uint32_t c0, uint32_t c1, uint32_t newid) {
    const int tid = threadIdx.x, nt = blockDim.x, G = cx.cfg.G;
    const bool self = c0 == c1;
    const int words = (p.n_out + 31) / 32;
    for (int r = 0; r < (self ? 1 : 2); ++r) {
        const uint32_t e = r == 0 ? c0 : c1;
        if ((int)(e % (uint32_t)G) != cx.rank)
            continue;
        const uint32_t off = ex.ws.cell_off[em_slot(ex.ws, e, G)], cnt = ex.ws.cell_cnt[em_slot(ex.ws, e, G)];
        uint2 *pl = ex.ws.cell_pl[ex.ws.ver[(size_t)cx.rank * ex.ws.e_cap + e] & 1]; // the version that has just become current
        for (uint32_t i = tid; i < cnt; i += nt)
            pl[off + i] = ex.D[r][ex.ws.cell_col[off + i]];
        for (int w = tid; w < words; w += nt)
            ex.ws.rowbits[em_slot(ex.ws, e, G) * ex.ws.words + w] = ex.B[r][w];
    }
    if ((int)(newid % (uint32_t)G) == cx.rank) {
        const int r = self ? 1 : 2;
        const int M = ex.eb->new_cells;
        const int off = ex.eb->pool_used; // (read by every thread before thread 0 advances it behind the barrier below)
        const bool fits = off + M <= ex.ws.pool_cap;
        const size_t pool = (size_t)cx.rank * ex.ws.pool_cap;
        if (fits)
            for (int o = tid; o < p.n_out; o += nt) {
                const uint2 v = ex.D[r][o];
                if ((v.x | v.y) != 0u) {
                    const int w = o >> 5;
                    const size_t ci = pool + off + ex.pre[w] + __popc(ex.B[r][w] & ((1u << (o & 31)) - 1u));
                    ex.ws.cell_col[ci] = (uint32_t)o;
                    ex.ws.cell_pl[0][ci] = v; // a new expression starts at version 0
                }
            }
        for (int w = tid; w < words; w += nt)
            ex.ws.rowbits[em_slot(ex.ws, newid, G) * ex.ws.words + w] = ex.B[r][w];
        __syncthreads(); // (uniform: the condition depends on newid and the CTA rank only)
        if (tid == 0) {
            ex.ws.cell_off[em_slot(ex.ws, newid, G)] = (uint32_t)((size_t)cx.rank * ex.ws.pool_cap + off);
            ex.ws.cell_cnt[em_slot(ex.ws, newid, G)] = fits ? (uint32_t)M : 0u;
            const int rn = ex.eb->n_mods - 1; // the new expression is the last rewritten row
            ex.ws.own_q[em_slot(ex.ws, newid, G)] = make_float4(ex.eb->mq[rn].min, ex.eb->mq[rn].max, ex.eb->mq[rn].step, ex.eb->ml[rn]);
            if (fits)
                ex.eb->pool_used = off + M;
            else
                cx.b->status = ST_LIST_OVERFLOW;
        }
    }
    __syncthreads();
}

// Pairs of one owned expression x with the rewritten rows flagged in `share` (dedup and ordering rules of
// state_opr.cc:307-340), executed by one warp: x's cells are read once, 32 per round (one per lane, coalesced; the first
// round arrives prefetched in c0 / col0); the cells that meet a digit of a rewritten row are broadcast one by one and
// every lane counts the pairs of ITS relative shift with popcounts.  Entries with count >= 2 go to this CTA's histogram
// segment.  Returns this lane's pairs.
template <int NR>
__device__ __forceinline__ int em_count_row(const ProblemDesc &p, const Ctx &cx, const EmCtx &ex, uint32_t x, uint32_t off, uint32_t cnt, uint32_t share, const QInt &qx, float lx,
                                            uint2 c0, uint32_t col0, uint32_t stamp, uint32_t thresh, Best &best) {
    const int lane = threadIdx.x & 31;
    const int nbits = p.nbits, n_sh = 2 * nbits - 1;
    const uint2 *plx = ex.ws.cell_pl[ex.ws.ver[(size_t)cx.rank * ex.ws.e_cap + x] & 1]; // (the owner's own cells)
    int pairs = 0;
    for (int s0 = 0; s0 < n_sh; s0 += 32) {
        const int si = s0 + lane, s = si - (nbits - 1);
        uint32_t same[NR], diff[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r)
            same[r] = diff[r] = 0u;
        for (uint32_t i0 = 0; i0 < cnt; i0 += 32) {
            uint2 c = c0;
            uint32_t col = col0;
            if (i0 > 0 || s0 > 0) { // (later rounds are fetched here)
                c = make_uint2(0u, 0u);
                col = 0u;
                if (i0 + lane < cnt) {
                    c = plx[off + i0 + lane];
                    col = ex.ws.cell_col[off + i0 + lane];
                }
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                if (!((share >> r) & 1u))
                    continue; // (uniform)
                const uint32_t m = ex.eb->mid[r];
                const bool self = x == m, x_first = x < m || self;
                uint2 v = make_uint2(0u, 0u);
                if ((c.x | c.y) != 0u)
                    v = ex.D[r][col];
                unsigned live = __ballot_sync(0xffffffffu, (v.x | v.y) != 0u);
                const bool active = si < n_sh && !(self && s >= 0);
                while (live) {
                    const int j = __ffs(live) - 1;
                    live &= live - 1;
                    const uint32_t cP = __shfl_sync(0xffffffffu, c.x, j), cN = __shfl_sync(0xffffffffu, c.y, j);
                    const uint32_t vP = __shfl_sync(0xffffffffu, v.x, j), vN = __shfl_sync(0xffffffffu, v.y, j);
                    if (!active)
                        continue;
                    const uint32_t Pl = x_first ? cP : vP, Nl = x_first ? cN : vN; // planes of the smaller id
                    const uint32_t Ph = x_first ? vP : cP, Nh = x_first ? vN : cN;
                    if (s >= 0) {
                        same[r] += __popc(Pl & (Ph >> s)) + __popc(Nl & (Nh >> s));
                        diff[r] += __popc(Pl & (Nh >> s)) + __popc(Nl & (Ph >> s));
                    }
                    else {
                        const int d = -s;
                        same[r] += __popc((Pl >> d) & Ph) + __popc((Nl >> d) & Nh);
                        diff[r] += __popc((Pl >> d) & Nh) + __popc((Nl >> d) & Ph);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            if (!((share >> r) & 1u))
                continue;
            pairs += (int)(same[r] + diff[r]);
            const uint32_t m = ex.eb->mid[r];
            const bool x_lo = x < m;
            const uint32_t lo = x_lo ? x : m, hi = x_lo ? m : x;
            const QInt qm = ex.eb->mq[r];
            const float lm = ex.eb->ml[r];
            if (same[r] >= 2u)
                emit_entry(p, cx, lo, hi, s, 0, same[r], x_lo ? qx : qm, x_lo ? lx : lm, x_lo ? qm : qx, x_lo ? lm : lx, stamp, thresh, best);
            if (diff[r] >= 2u)
                emit_entry(p, cx, lo, hi, s, 1, diff[r], x_lo ? qx : qm, x_lo ? lx : lm, x_lo ? qm : qx, x_lo ? lm : lx, stamp, thresh, best);
        }
    }
    return pairs;
}

// Dense variant for one rewritten row r: lanes are CELLS (32 columns per round), every lane keeps a private counter per
// relative shift (|shift| < MAXB, compile-time unrolled -- planes have no bits at or above nbits, so shifts beyond the
// CSD width count nothing) and the counters are summed over the warp at the end.  About three times fewer instructions
// per digit pair than broadcasting cell by cell when most columns hold digits of both rows (the early steps).
template <int MAXB>
__device__ __noinline__ int em_count_dense(const ProblemDesc &p, const Ctx &cx, const EmCtx &ex, uint32_t x, uint32_t off, uint32_t cnt, int r, const QInt &qx, float lx, uint32_t stamp,
                                            uint32_t thresh, Best &best) {
    constexpr int NS = 2 * MAXB - 1;
    const int lane = threadIdx.x & 31;
    const uint2 *plx = ex.ws.cell_pl[ex.ws.ver[(size_t)cx.rank * ex.ws.e_cap + x] & 1];
    const uint2 *Dm = ex.D[r];
    const uint32_t m = ex.eb->mid[r];
    const bool self = x == m, x_first = x < m || self;
    uint32_t same[NS], diff[NS];
#pragma unroll
    for (int si = 0; si < NS; ++si)
        same[si] = diff[si] = 0u;
    for (uint32_t i0 = 0; i0 < cnt; i0 += 32) {
        uint2 c = make_uint2(0u, 0u), v = make_uint2(0u, 0u);
        if (i0 + lane < cnt) {
            c = plx[off + i0 + lane];
            if ((c.x | c.y) != 0u)
                v = Dm[ex.ws.cell_col[off + i0 + lane]];
        }
        if ((v.x | v.y) == 0u)
            continue;
        const uint32_t Pl = x_first ? c.x : v.x, Nl = x_first ? c.y : v.y; // planes of the smaller id
        const uint32_t Ph = x_first ? v.x : c.x, Nh = x_first ? v.y : c.y;
#pragma unroll
        for (int si = 0; si < NS; ++si) {
            const int sh = si - (MAXB - 1);
            if (sh >= 0) {
                if (!self) {
                    same[si] += __popc(Pl & (Ph >> sh)) + __popc(Nl & (Nh >> sh));
                    diff[si] += __popc(Pl & (Nh >> sh)) + __popc(Nl & (Ph >> sh));
                }
            }
            else {
                same[si] += __popc((Pl >> -sh) & Ph) + __popc((Nl >> -sh) & Nh);
                diff[si] += __popc((Pl >> -sh) & Nh) + __popc((Nl >> -sh) & Ph);
            }
        }
    }
    uint32_t my_same = 0u, my_diff = 0u; // lane si ends up with the totals of shift si - (MAXB - 1)
#pragma unroll
    for (int si = 0; si < NS; ++si) {
        uint32_t a = same[si], d = diff[si];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        if (lane == si) {
            my_same = a;
            my_diff = d;
        }
    }
    const int sh = lane - (MAXB - 1);
    const bool x_lo = x < m;
    const uint32_t lo = x_lo ? x : m, hi = x_lo ? m : x;
    const QInt qm = ex.eb->mq[r];
    const float lm = ex.eb->ml[r];
    if (my_same >= 2u)
        emit_entry(p, cx, lo, hi, sh, 0, my_same, x_lo ? qx : qm, x_lo ? lx : lm, x_lo ? qm : qx, x_lo ? lm : lx, stamp, thresh, best);
    if (my_diff >= 2u)
        emit_entry(p, cx, lo, hi, sh, 1, my_diff, x_lo ? qx : qm, x_lo ? lx : lm, x_lo ? qm : qx, x_lo ? lm : lx, stamp, thresh, best);
    return (int)(my_same + my_diff);
}
#ifndef DA_EM_DENSE_CELLS
#define DA_EM_DENSE_CELLS 48 // rows with at least this many cells take the dense variant (CSD width <= 9 bits, NS <= 32)
#endif

// C. recount of everything this CTA owns: tiles of blockDim.x owned expressions; a thread per expression tests its
// column bitmap against the rewritten rows and fetches what the counting needs (all loads of a tile in flight together),
// then one warp per active expression counts, with the first cells of its next expression already on their way
__device__ void em_recount(const ProblemDesc &p, const Ctx &cx, const EmCtx &ex, uint32_t newid, uint32_t stamp, uint32_t thresh, Best &best) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5, G = cx.cfg.G;
    const int words = (p.n_out + 31) / 32;
    const int n_owned = (int)newid >= cx.rank ? ((int)newid - cx.rank) / G + 1 : 0;
    const int n_mods = ex.eb->n_mods;
    const unsigned char *ver = ex.ws.ver + (size_t)cx.rank * ex.ws.e_cap;
    int nr = 0;
    for (int base = 0; base < n_owned; base += nt) { // (uniform bounds)
        const int i = base + tid;
        const uint32_t x = (uint32_t)(cx.rank + G * i);
        uint32_t share = 0u; // bit r: x has a digit in a column where rewritten row r has one
        uint32_t xoff = 0u, xcnt = 0u;
        QInt qx;
        float lx = 0.0f;
        qx.min = qx.max = qx.step = 0.0f;
        if (i < n_owned) {
            const size_t slot = (size_t)cx.rank * ex.ws.per + (size_t)i; // == em_slot(x): consecutive threads, consecutive slots
            const uint32_t *rb = ex.ws.rowbits + slot * ex.ws.words;
            xoff = ex.ws.cell_off[slot];
            xcnt = ex.ws.cell_cnt[slot];
            bool xmod = false;
            for (int r = 0; r < n_mods; ++r)
                xmod = xmod || ex.eb->mid[r] == x;
            const float4 oq = ex.ws.own_q[slot]; // (records never change; the new expression's was stored by em_update_owned)
            qx.min = oq.x, qx.max = oq.y, qx.step = oq.z, lx = oq.w;
            for (int w = 0; w < words; ++w) {
                const uint32_t v = rb[w];
                for (int r = 0; r < n_mods; ++r)
                    share |= (v & ex.B[r][w]) != 0u ? 1u << r : 0u;
            }
            for (int r = 0; r < n_mods; ++r)
                if (xmod && ex.eb->mid[r] > x)
                    share &= ~(1u << r); // pairs among the rewritten rows are counted once, at the larger id
        }
        const bool on = share != 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        int wbase = 0;
        if (lane == 0 && bal)
            wbase = smem_add(&ex.eb->tile_n, __popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (on) {
            const int k = wbase + __popc(bal & ((1u << lane) - 1u));
            ex.tile[k] = x;
            ex.tile_off[k] = xoff;
            ex.tile_cnt[k] = xcnt | (share << 28);
            ex.tile_q[k] = make_float4(qx.min, qx.max, qx.step, lx);
        }
        __syncthreads();
        const int n_tile = ex.eb->tile_n;
        // software pipeline over this warp's expressions: the first 32 cells of the next one are requested before the
        // current one is counted
        uint2 c_nx = make_uint2(0u, 0u);
        uint32_t col_nx = 0u;
        if (wid < n_tile) {
            const uint32_t xx = ex.tile[wid], off = ex.tile_off[wid], cnt = ex.tile_cnt[wid] & 0x0fffffffu;
            if ((uint32_t)lane < cnt) {
                c_nx = ex.ws.cell_pl[ver[xx] & 1][off + lane];
                col_nx = ex.ws.cell_col[off + lane];
            }
        }
        for (int k = wid; k < n_tile; k += nw) {
            const uint32_t xx = ex.tile[k], off = ex.tile_off[k], cs = ex.tile_cnt[k];
            const float4 q4 = ex.tile_q[k];
            const uint2 c_cur = c_nx;
            const uint32_t col_cur = col_nx;
            c_nx = make_uint2(0u, 0u);
            col_nx = 0u;
            if (k + nw < n_tile) {
                const uint32_t xn = ex.tile[k + nw], offn = ex.tile_off[k + nw], cntn = ex.tile_cnt[k + nw] & 0x0fffffffu;
                if ((uint32_t)lane < cntn) {
                    c_nx = ex.ws.cell_pl[ver[xn] & 1][offn + lane];
                    col_nx = ex.ws.cell_col[offn + lane];
                }
            }
            QInt qq;
            qq.min = q4.x, qq.max = q4.y, qq.step = q4.z;
            const uint32_t cnt = cs & 0x0fffffffu;
            if (cnt >= DA_EM_DENSE_CELLS && p.nbits <= 9) {
                for (int r = 0; r < n_mods; ++r)
                    if ((cs >> (28 + r)) & 1u)
                        nr += em_count_dense<9>(p, cx, ex, xx, off, cnt, r, qq, q4.w, stamp, thresh, best);
            }
            else
                nr += em_count_row<3>(p, cx, ex, xx, off, cnt, cs >> 28, qq, q4.w, c_cur, col_cur, stamp, thresh, best);
        }
        __syncthreads();
        if (tid == 0)
            ex.eb->tile_n = 0;
        __syncthreads();
    }
    // (em_count_row returns per-lane counts: lanes are shifts)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        nr += __shfl_xor_sync(0xffffffffu, nr, off);
    if (lane == 0 && nr)
        smem_add(&cx.b->r_step, nr);
}

// D. before the adder trees: scatter the live cells into per-column lists in global memory (group-wide)
__device__ void em_scatter_columns(const ProblemDesc &p, const Ctx &cx, const EmCtx &ex, uint32_t n_expr) {
    const int tid = threadIdx.x, nt = blockDim.x, G = cx.cfg.G;
    for (int o = cx.rank * nt + tid; o < p.n_out; o += G * nt)
        cx.ws.col_len[o] = 0;
    group_sync(cx);
    for (uint32_t x = (uint32_t)cx.rank; x < n_expr; x += (uint32_t)G) {
        const uint32_t off = ex.ws.cell_off[em_slot(ex.ws, x, G)], cnt = ex.ws.cell_cnt[em_slot(ex.ws, x, G)];
        const uint2 *pl = ex.ws.cell_pl[ex.ws.ver[(size_t)cx.rank * ex.ws.e_cap + x] & 1];
        for (uint32_t i = tid; i < cnt; i += nt) {
            const uint2 c = pl[off + i];
            if ((c.x | c.y) == 0u)
                continue;
            const uint32_t o = ex.ws.cell_col[off + i];
            const int pos = atomicAdd(&cx.ws.col_len[o], 1);
            if (pos >= p.col_cap) {
                cx.b->status = ST_LIST_OVERFLOW;
                continue;
            }
            uint32_t *col = cx.ws.col_u32 + (size_t)o * 3 * p.col_cap;
            col[pos] = x;
            col[p.col_cap + pos] = c.x;
            col[2 * p.col_cap + pos] = c.y;
        }
    }
    group_sync(cx);
}

} // namespace da
