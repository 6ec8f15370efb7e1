// CPU model (not product, not tests) of an expression-major organisation of the greedy step, checked against the
// oracle restatement at every step:
//   * a dense row table planes[e][o] = (P, N) sign planes of expression e in output column o;
//   * expression e is owned by CTA e % G; the owner recounts every pair (m, e), m in {c0, c1, new}, over ALL columns
//     with bit-parallel popcounts and emits the histogram entries locally -- no cross-CTA counters, no harvest;
//   * the substitution is computed redundantly by everybody from the rows of c0 and c1.
// build: g++ -O2 -std=c++17 -fopenmp scripts/proto_expr_major.cc -o /tmp/proto; run: /tmp/proto 32 8 [G]
#include "../oracle/cmvm_oracle.cc"
#include <random>
using namespace orc;
typedef std::pair<uint32_t, uint32_t> PN;

static PN planes_of(const Row &r) {
    uint32_t P = 0, N = 0;
    for (int8_t v : r)
        (v > 0 ? P : N) |= 1u << dshift(v);
    return {P, N};
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 32, bits = argc > 2 ? atoi(argv[2]) : 8, G = argc > 3 ? atoi(argv[3]) : 14;
    const std::string method = argc > 4 ? argv[4] : "wmc";
    std::mt19937 rng(1);
    std::uniform_int_distribution<int> d(-(1 << (bits - 1)), (1 << (bits - 1)) - 1);
    Mat k((size_t)n * n);
    for (auto &v : k) v = (float)d(rng);
    std::vector<QI> q(n, QI{-128, 127, 1});
    std::vector<float> lat(n, 0.f);
    Stage cnt;
    State st = create_state(k, n, n, q, lat, false, &cnt);
    const int n_out = n, NB = st.n_bits;
    std::vector<std::vector<PN>> planes(n, std::vector<PN>(n_out));
    for (int e = 0; e < n; ++e)
        for (int o = 0; o < n_out; ++o)
            planes[e][o] = planes_of(st.expr[e][o]);
    std::map<Key, uint32_t> H = st.freq; // the initial histogram is built by the (unchanged) popcount kernel
    long long T = 0, tests = 0, active_expr = 0, active_cells = 0, all_cells = 0, emitted = 0;
    std::vector<long long> per_cta(G);
    double imbalance = 0;
    while (!st.freq.empty()) {
        Key key;
        if (!select_pair(st, method, key)) break;
        const int64_t c0 = key.id0, c1 = key.id1;
        const int shift = key.shift;
        const bool sub = key.sub;
        // ---- substitution from the two rows (every CTA, redundantly); same plane algebra as the CUDA column_substitute
        const int64_t newid = (int64_t)planes.size();
        std::vector<PN> r0 = planes[c0], r1 = planes[c1], rn(n_out, PN{0, 0});
        for (int o = 0; o < n_out; ++o) {
            uint32_t P0 = r0[o].first, N0 = r0[o].second, P1 = r1[o].first, N1 = r1[o].second, Pn = 0, Nn = 0;
            if (c0 != c1) {
                const bool flip = shift < 0;
                const int rel = flip ? -shift : shift;
                const uint32_t AP = flip ? P1 : P0, AN = flip ? N1 : N0, BP = flip ? P0 : P1, BN = flip ? N0 : N1;
                const uint32_t M = sub ? ((AP & (BN >> rel)) | (AN & (BP >> rel))) : ((AP & (BP >> rel)) | (AN & (BN >> rel)));
                const uint32_t MB = M << rel;
                if (!flip) { Pn = AP & M; Nn = AN & M; P0 = AP & ~M; N0 = AN & ~M; P1 = BP & ~MB; N1 = BN & ~MB; }
                else { Pn = BP & MB; Nn = BN & MB; P1 = AP & ~M; N1 = AN & ~M; P0 = BP & ~MB; N0 = BN & ~MB; }
            }
            else {
                const int rel = -shift;
                const uint32_t live = P0 | N0;
                uint32_t tomb = 0;
                for (uint32_t m = live; m; m &= m - 1) {
                    const int pl = __builtin_ctz(m);
                    if ((tomb >> pl) & 1) continue;
                    const int qq = pl + rel;
                    if (qq >= NB || qq >= 32) continue;
                    if (!((live >> qq) & 1) || ((tomb >> qq) & 1)) continue;
                    if ((((N0 >> pl) ^ (N0 >> qq)) & 1) != (uint32_t)sub) continue;
                    if ((N0 >> qq) & 1) Nn |= 1u << qq; else Pn |= 1u << qq;
                    tomb |= (1u << pl) | (1u << qq);
                }
                P0 &= ~tomb; N0 &= ~tomb; P1 = P0; N1 = N0;
            }
            r0[o] = {P0, N0}; r1[o] = {P1, N1}; rn[o] = {Pn, Nn};
        }
        planes[c0] = r0;
        planes[c1] = r1;
        planes.push_back(rn);
        // ---- oracle step
        substitute(st, key, -1, -1);
        recount(st, key);
        for (int64_t e : {c0, c1, newid})
            for (int o = 0; o < n_out; ++o)
                if (planes_of(st.expr[e][o]) != planes[e][o]) { printf("substitution mismatch step %lld expr %lld col %d\n", T, (long long)e, o); return 1; }
        // ---- lazy purge of everything touching c0 / c1
        for (auto it = H.begin(); it != H.end();)
            it = (it->first.id0 == c0 || it->first.id0 == c1 || it->first.id1 == c0 || it->first.id1 == c1) ? H.erase(it) : std::next(it);
        // ---- owners recount
        uint32_t act[64] = {0}; // active-column bitmap of the three rewritten rows
        for (int o = 0; o < n_out; ++o)
            if ((r0[o].first | r0[o].second | r1[o].first | r1[o].second | rn[o].first | rn[o].second) != 0) act[o >> 5] |= 1u << (o & 31);
        std::vector<int64_t> mods = {c0};
        if (c1 != c0) mods.push_back(c1);
        mods.push_back(newid);
        std::fill(per_cta.begin(), per_cta.end(), 0);
        for (int r = 0; r < G; ++r)
            for (int64_t x = r; x <= newid; x += G) {
                ++tests;
                bool any = false;
                int cells = 0;
                for (int o = 0; o < n_out; ++o)
                    if ((planes[x][o].first | planes[x][o].second) != 0) { ++cells; if ((act[o >> 5] >> (o & 31)) & 1) any = true; }
                all_cells += cells;
                if (!any) continue;
                ++active_expr;
                const bool xmod = x == c0 || x == c1 || x == newid;
                for (int64_t m : mods) {
                    if (xmod && m > x) continue; // dedup rule, state_opr.cc:310-312
                    const int64_t lo = std::min(m, x), hi = std::max(m, x);
                    const std::vector<PN> &L = planes[lo], &Hh = planes[hi];
                    uint32_t same[64] = {0}, diff[64] = {0}; // index shift + NB - 1
                    for (int o = 0; o < n_out; ++o) {
                        const uint32_t Pl = L[o].first, Nl = L[o].second, Ph = Hh[o].first, Nh = Hh[o].second;
                        if ((Pl | Nl) == 0 || (Ph | Nh) == 0) continue;
                        ++active_cells; ++per_cta[r];
                        for (int s = -(NB - 1); s <= NB - 1; ++s) {
                            if (lo == hi && s >= 0) continue;
                            uint32_t sm, df;
                            if (s >= 0) { sm = __builtin_popcount(Pl & (Ph >> s)) + __builtin_popcount(Nl & (Nh >> s)); df = __builtin_popcount(Pl & (Nh >> s)) + __builtin_popcount(Nl & (Ph >> s)); }
                            else { const int dd = -s; sm = __builtin_popcount((Pl >> dd) & Ph) + __builtin_popcount((Nl >> dd) & Nh); df = __builtin_popcount((Pl >> dd) & Nh) + __builtin_popcount((Nl >> dd) & Ph); }
                            same[s + NB - 1] += sm; diff[s + NB - 1] += df;
                        }
                    }
                    for (int s = -(NB - 1); s <= NB - 1; ++s) {
                        if (same[s + NB - 1] >= 2) { H[Key{hi, lo, false, (int8_t)s}] = same[s + NB - 1]; ++emitted; }
                        if (diff[s + NB - 1] >= 2) { H[Key{hi, lo, true, (int8_t)s}] = diff[s + NB - 1]; ++emitted; }
                    }
                }
            }
        const long long mx = *std::max_element(per_cta.begin(), per_cta.end());
        long long sm = 0;
        for (long long v : per_cta) sm += v;
        if (sm) imbalance += (double)mx * G / sm;
        // ---- compare with the oracle's histogram
        if (H.size() != st.freq.size() || !std::equal(H.begin(), H.end(), st.freq.begin(), [](auto &a, auto &b) { return !(a.first < b.first) && !(b.first < a.first) && a.second == b.second; })) {
            printf("histogram mismatch at step %lld: %zu vs %zu entries\n", T, H.size(), st.freq.size());
            return 1;
        }
        ++T;
    }
    printf("OK n=%d bits=%d G=%d method=%s: %lld steps identical (planes + histogram). per step: %.0f owner tests, %.1f active expressions, %.0f active (expr,col) cells of %.0f, %.1f entries emitted; max/mean CTA load %.2f\n",
           n, bits, G, method.c_str(), T, (double)tests / T, (double)active_expr / T, (double)active_cells / T, (double)all_cells / T, (double)emitted / T, imbalance / T);
}
