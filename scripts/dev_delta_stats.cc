// dev analysis (not product, not tests): work an incremental histogram update would do per greedy step, vs the full recount
// build: g++ -O2 -std=c++17 -fopenmp scripts/dev_delta_stats.cc -o /tmp/delta_stats; run: /tmp/delta_stats 64 8
#include "../oracle/cmvm_oracle.cc"
#include <random>
using namespace orc;
int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 64, bits = argc > 2 ? atoi(argv[2]) : 8;
    std::mt19937 rng(0);
    std::uniform_int_distribution<int> d(-(1 << (bits - 1)), (1 << (bits - 1)) - 1);
    Mat k((size_t)n * n);
    for (auto &v : k) v = (float)d(rng);
    std::vector<QI> q(n, QI{-128, 127, 1});
    std::vector<float> lat(n, 0.f);
    Stage cnt;
    State st = create_state(k, n, n, q, lat, false, &cnt);
    long long T = 0, sumR = 0, sumD = 0, sumN = 0, sumNewEnt = 0, sumChanged = 0, sumF = 0, sumOldEnt = 0;
    long long F0 = st.freq.size(), Fmax = F0;
    std::vector<long long> coldig(n);
    while (!st.freq.empty()) {
        Key key;
        if (!select_pair(st, "wmc", key)) break;
        sumF += st.freq.size();
        // snapshot
        std::vector<Row> o0 = st.expr[key.id0], o1 = st.expr[key.id1];
        std::map<Key, uint32_t> old;
        for (auto &kv : st.freq) {
            const Key &kk = kv.first;
            if (kk.id0 == key.id0 || kk.id0 == key.id1 || kk.id1 == key.id0 || kk.id1 == key.id1) old.insert(kv);
        }
        substitute(st, key, -1, -1);
        long long R = recount(st, key);
        sumR += R;
        long long newid = st.expr.size() - 1;
        long long D = 0, N = 0;
        for (int o = 0; o < n; ++o) {
            long long lost = (long long)o0[o].size() - (long long)st.expr[key.id0][o].size();
            if (key.id0 != key.id1) lost += (long long)o1[o].size() - (long long)st.expr[key.id1][o].size();
            long long nd = st.expr[newid][o].size();
            if (lost == 0 && nd == 0) continue;
            long long tot = 0;
            for (size_t x = 0; x < st.expr.size(); ++x) tot += st.expr[x][o].size();
            D += lost * tot;
            N += nd * tot;
        }
        sumD += D; sumN += N;
        long long ne = 0, changed = 0;
        for (auto &kv : st.freq) if (kv.first.id1 == newid) ++ne;
        for (auto &kv : old) {
            auto it = st.freq.find(kv.first);
            if (it == st.freq.end() || it->second != kv.second) ++changed;
        }
        sumNewEnt += ne; sumChanged += changed; sumOldEnt += old.size();
        Fmax = std::max<long long>(Fmax, st.freq.size());
        if (T < 5 || T % 2000 == 0) printf("t=%lld F=%zu R=%lld D=%lld N=%lld newEnt=%lld changed=%lld oldEnt=%zu\n", T, st.freq.size(), R, D, N, ne, changed, old.size());
        ++T;
    }
    printf("n=%d bits=%d T=%lld F0=%lld Fmax=%lld sumF=%lld sumR=%lld sumDelta=%lld sumNewPairs=%lld sumNewEnt=%lld sumChanged=%lld sumOldEnt=%lld\n", n, bits, T, F0, Fmax, sumF, sumR, sumD, sumN, sumNewEnt, sumChanged, sumOldEnt);
}
