// oracle/cmvm_oracle.cc -- CPU restatement of da4ml's CMVM solver.  TEST INFRASTRUCTURE ONLY.
//
// A from-scratch, dependency-free C++17 restatement of the reference algorithm under
// /root/reference/src/da4ml/_binary/cmvm/ (file:line cited at every function).  It is the checker
// that travels with the repository (the GPU box has no /root/reference); it is pinned against the
// reference's own object code (oracle/_ref, built from the unmodified reference TUs) by
// tests/test_oracle.py and against the committed golden vectors in tests/golden/.
// Deliberately written with ordinary containers (std::map histogram, per-row digit vectors) and
// plain libm calls, i.e. neither the reference's flat-vector code nor the product's bit-plane /
// atomic-counter formulation, so that agreement between the three is meaningful.
//
// Never imported, linked or executed by the product (da4ml_b200/); only tests/, smoke() and
// bench.py's CPU-baseline legs use it.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

struct QI {
    float min, max, step;
};
struct OpRec {
    int64_t id0, id1, opcode, data;
    QI q;
    float lat, cost;
};
struct Stage {
    int64_t n_in = 0, n_out = 0;
    std::vector<int64_t> inp_shifts, out_idxs, out_shifts, out_negs;
    std::vector<OpRec> ops;
    int carry_size = -1, adder_size = -1;
    // work counters
    int64_t T = 0, sum_f = 0, sum_r = 0, f0 = 0, r0 = 0, d0 = 0, d_final = 0;
};
using Mat = std::vector<float>; // row-major

// ---------------------------------------------------------------- bit_decompose.cc:10-20
static int lsb_loc(float x) {
    if (x == 0.0f)
        return 127;
    uint32_t bits;
    std::memcpy(&bits, &x, 4);
    int exp = (bits >> 23) & 0xFF;
    uint32_t mant = bits & 0x7FFFFF;
    return (int8_t)(exp + __builtin_ctz(mant + (1u << 23)) - 150);
}
// indexers.hh:12-18
static int8_t iceil_log2(float x) {
    uint32_t bits;
    std::memcpy(&bits, &x, 4);
    uint8_t exp = (bits >> 23) & 0xFF;
    return (int8_t)(exp - 127 + ((bits & 0x7FFFFF) != 0));
}

// bit_decompose.hh:25-34: columns first, then rows of the column-scaled matrix
static void center(Mat &a, int n_in, int n_out, std::vector<int8_t> &s0, std::vector<int8_t> &s1) {
    s0.assign(n_in, 0);
    s1.assign(n_out, 0);
    for (int j = 0; j < n_out; ++j) {
        int m = 127;
        for (int i = 0; i < n_in; ++i)
            m = std::min(m, lsb_loc(a[(size_t)i * n_out + j]));
        s1[j] = (int8_t)m;
    }
    for (int i = 0; i < n_in; ++i)
        for (int j = 0; j < n_out; ++j)
            a[(size_t)i * n_out + j] = (float)(a[(size_t)i * n_out + j] * std::pow(2.0, -(int)s1[j]));
    for (int i = 0; i < n_in; ++i) {
        int m = 127;
        for (int j = 0; j < n_out; ++j)
            m = std::min(m, lsb_loc(a[(size_t)i * n_out + j]));
        s0[i] = (int8_t)m;
    }
    for (int i = 0; i < n_in; ++i)
        for (int j = 0; j < n_out; ++j)
            a[(size_t)i * n_out + j] = (float)(a[(size_t)i * n_out + j] * std::pow(2.0, -(int)s0[i]));
}

// bit_decompose.cc:22-42: digits[e*N + n] in {-1,0,1}; returns N
static int to_csd(std::vector<int32_t> x, std::vector<int8_t> &digits) {
    int32_t mx = 0;
    for (int32_t v : x)
        mx = std::max(mx, (int32_t)std::abs(v));
    size_t N = (size_t)std::ceil(std::log2(std::max((float)mx, 1.0f) * 1.5));
    N = std::max(N, (size_t)1);
    digits.assign(x.size() * N, 0);
    for (int n = (int)N - 1; n >= 0; --n) {
        int32_t p2 = (int32_t)(1U << n);
        int32_t thres = p2 * 2 / 3;
        for (size_t e = 0; e < x.size(); ++e) {
            int d = (x[e] > thres) - (x[e] < -thres);
            digits[e * N + n] = (int8_t)d;
            x[e] -= p2 * d;
        }
    }
    return (int)N;
}

// bit_decompose.cc:44-62
static int csd_decompose(const Mat &k, int n_in, int n_out, bool do_center, std::vector<int8_t> &digits, std::vector<int8_t> &s0, std::vector<int8_t> &s1) {
    Mat a(k);
    s0.assign(n_in, 0);
    s1.assign(n_out, 0);
    if (do_center)
        center(a, n_in, n_out, s0, s1);
    std::vector<int32_t> xi(a.size());
    for (size_t e = 0; e < a.size(); ++e)
        xi[e] = (int32_t)a[e];
    return to_csd(xi, digits);
}

// ---------------------------------------------------------------- state_opr.cc:8-67
static QI qint_add(QI q0, QI q1, int64_t shift, bool sub0, bool sub1) {
    if (sub0) {
        std::swap(q0.min, q0.max);
        q0.min = -q0.min;
        q0.max = -q0.max;
    }
    if (sub1) {
        std::swap(q1.min, q1.max);
        q1.min = -q1.min;
        q1.max = -q1.max;
    }
    float s = std::pow(2.0, shift);
    q1.min *= s;
    q1.max *= s;
    q1.step *= s;
    return QI{q0.min + q1.min, q0.max + q1.max, std::min(q0.step, q1.step)};
}
static std::pair<float, float> cost_add(QI q0, QI q1, int64_t shift, bool sub, int adder_size, int carry_size) {
    if (adder_size < 0 && carry_size < 0)
        return {1.0f, 1.0f};
    if (adder_size < 0)
        adder_size = 65535;
    if (carry_size < 0)
        carry_size = 65535;
    float min0 = q0.min, max0 = q0.max, step0 = q0.step;
    float min1 = q1.min, max1 = q1.max, step1 = q1.step;
    if (sub)
        std::swap(min1, max1);
    float sf = std::pow(2.0, shift);
    min1 *= sf;
    max1 *= sf;
    step1 *= sf;
    max0 += step0;
    max1 += step1;
    float f = -std::log2(std::max(step0, step1));
    float i = std::ceil(std::log2(std::max({std::abs(min0), std::abs(min1), std::abs(max0), std::abs(max1)})));
    int k = (q0.min < 0 || q1.min < 0) ? 1 : 0;
    float n_accum = k + i + f;
    return {std::ceil(n_accum / carry_size), std::ceil(n_accum / adder_size)};
}
// indexers.cc:36-56
static std::pair<int8_t, int8_t> overlap_and_accum(QI q0, QI q1) {
    float max0 = q0.max + q0.step, max1 = q1.max + q1.step;
    int8_t f = -iceil_log2(std::max(q0.step, q1.step));
    int8_t i_high = iceil_log2((float)std::max({std::abs(q0.min), std::abs(q1.min), std::abs(max0), std::abs(max1)}));
    int8_t i_low = iceil_log2(std::min(std::max(std::abs(q0.min), std::abs(max0)), std::max(std::abs(q1.min), std::abs(max1))));
    int8_t k = (q0.min < 0 || q1.min < 0) ? 1 : 0;
    int8_t n_accum = k + i_high + f;
    int8_t n_overlap = k + i_low + f;
    return {n_overlap, n_accum};
}

// ---------------------------------------------------------------- the greedy state
// A digit is stored as the reference does: +/-(shift+1) (types.hh:104-141); a row is the ascending
// list of the digits of one expression in one output column.
using Row = std::vector<int8_t>;
struct Key { // ordering == Pair::operator< (types.hh:28-36)
    int64_t id1, id0;
    bool sub;
    int8_t shift;
    bool operator<(const Key &o) const { return std::tie(id1, id0, sub, shift) < std::tie(o.id1, o.id0, o.sub, o.shift); }
};
struct State {
    int n_in = 0, n_out = 0, n_bits = 0;
    std::vector<int8_t> shift0, shift1;
    std::vector<std::vector<Row>> expr; // expr[e][o]
    std::vector<OpRec> ops;
    std::map<Key, uint32_t> freq; // only entries with count >= 2 (types.hh:83)
};
static inline int dshift(int8_t v) { return std::abs(v) - 1; }
static inline int dsign(int8_t v) { return v > 0 ? 1 : -1; }
static Key make_key(int64_t lo, int64_t hi, int8_t vlo, int8_t vhi) { // state_opr.cc:69-77
    if (lo > hi)
        throw std::invalid_argument("id0 should be <= id1");
    return Key{hi, lo, dsign(vlo) != dsign(vhi), (int8_t)(dshift(vhi) - dshift(vlo))};
}
// all digit pairs of (lo, hi) in column o appended to raw (state_opr.cc:122-138 / :320-337)
static void pairs_of(const State &st, int64_t lo, int64_t hi, int o, std::vector<Key> &raw) {
    const Row &rl = st.expr[lo][o], &rh = st.expr[hi][o];
    if (rl.empty() || rh.empty())
        return;
    if (lo == hi) {
        for (size_t a = 1; a < rl.size(); ++a)
            for (size_t b = 0; b < a; ++b)
                raw.push_back(make_key(lo, lo, rl[a], rl[b]));
    }
    else {
        for (int8_t v0 : rl)
            for (int8_t v1 : rh)
                raw.push_back(make_key(lo, hi, v0, v1));
    }
}
static int64_t add_counted(State &st, std::vector<Key> &raw) { // FreqMap::batch_add types.hh:73-95
    std::map<Key, uint32_t> cnt;
    for (const Key &k : raw)
        cnt[k]++;
    for (auto &kv : cnt)
        if (kv.second >= 2)
            st.freq[kv.first] = kv.second;
    return (int64_t)raw.size();
}

// state_opr.cc:79-159
static State create_state(const Mat &kernel, int n_in, int n_out, const std::vector<QI> &q, const std::vector<float> &lat, bool no_stat, Stage *cnt) {
    State st;
    st.n_in = n_in;
    st.n_out = n_out;
    std::vector<int8_t> digits;
    st.n_bits = csd_decompose(kernel, n_in, n_out, true, digits, st.shift0, st.shift1);
    const int N = st.n_bits;
    st.expr.assign(n_in, std::vector<Row>(n_out));
    for (int i = 0; i < n_in; ++i) {
        bool dead = q[i].min == 0.0 && q[i].max == 0.0; // state_opr.cc:92-97
        for (int o = 0; o < n_out; ++o)
            for (int j = 0; j < N; ++j) {
                int8_t v = dead ? 0 : digits[((size_t)i * n_out + o) * N + j];
                if (v != 0)
                    st.expr[i][o].push_back((int8_t)(v * (j + 1)));
            }
    }
    if (!no_stat) {
        std::vector<Key> raw;
        for (int o = 0; o < n_out; ++o)
            for (int i0 = 0; i0 < n_in; ++i0)
                for (int i1 = i0; i1 < n_in; ++i1)
                    pairs_of(st, i0, i1, o, raw);
        int64_t r = add_counted(st, raw);
        if (cnt) {
            cnt->r0 = r;
            cnt->f0 = (int64_t)st.freq.size();
        }
    }
    for (int i = 0; i < n_in; ++i)
        st.ops.push_back(OpRec{i, -1, -1, 0, q[i], lat[i], 0.0f});
    if (cnt)
        for (auto &e : st.expr)
            for (auto &r : e)
                cnt->d0 += (int64_t)r.size();
    return st;
}

// ---------------------------------------------------------------- selectors, indexers.cc:6-90
// Iterate in key order; ">=" keeps the LAST maximum.  Returns false when nothing qualifies.
static bool select_pair(const State &st, const std::string &m, Key &out) {
    bool found = false;
    if (m == "mc") {
        size_t best = 0;
        for (auto &kv : st.freq)
            if (kv.second >= best) {
                best = kv.second;
                out = kv.first;
                found = true;
            }
        return found;
    }
    if (m == "mc-dc" || m == "mc-pdc") {
        float best = (m == "mc-dc") ? 0 : -std::numeric_limits<float>::infinity();
        float factor = 1e9;
        for (auto &kv : st.freq) {
            float l0 = st.ops[kv.first.id0].lat, l1 = st.ops[kv.first.id1].lat;
            float score = kv.second - factor * std::abs(l0 - l1);
            if (score >= best) {
                best = score;
                out = kv.first;
                found = true;
            }
        }
        return found;
    }
    if (m == "wmc") {
        int64_t best = 0;
        for (auto &kv : st.freq) {
            auto [ov, ac] = overlap_and_accum(st.ops[kv.first.id0].q, st.ops[kv.first.id1].q);
            (void)ac;
            int64_t score = int64_t(kv.second) * ov;
            if (score >= best) {
                best = score;
                out = kv.first;
                found = true;
            }
        }
        return found;
    }
    if (m == "wmc-dc" || m == "wmc-pdc") {
        float best = (m == "wmc-dc") ? 0 : -std::numeric_limits<float>::infinity();
        for (auto &kv : st.freq) {
            auto [ov, ac] = overlap_and_accum(st.ops[kv.first.id0].q, st.ops[kv.first.id1].q);
            (void)ac;
            float l0 = st.ops[kv.first.id0].lat, l1 = st.ops[kv.first.id1].lat;
            float score = kv.second * ov - 256 * std::abs(l0 - l1); // uint32 * int8: unsigned arithmetic, as in the reference
            if (score >= best) {
                best = score;
                out = kv.first;
                found = true;
            }
        }
        return found;
    }
    throw std::runtime_error("Unknown method: " + m);
}

// state_opr.cc:227-283
static void substitute(State &st, const Key &pr, int adder_size, int carry_size) {
    const int64_t pid0 = pr.id0, pid1 = pr.id1;
    auto [dlat, cost] = cost_add(st.ops[pid0].q, st.ops[pid1].q, pr.shift, pr.sub, adder_size, carry_size);
    float lat = std::max(st.ops[pid0].lat, st.ops[pid1].lat) + dlat;
    QI q = qint_add(st.ops[pid0].q, st.ops[pid1].q, pr.shift, false, pr.sub);
    st.ops.push_back(OpRec{pid0, pid1, (int64_t)pr.sub, pr.shift, q, lat, cost});

    int64_t a = pid0, b = pid1;
    int rel = pr.shift;
    bool flip = false;
    if (rel < 0) {
        std::swap(a, b);
        rel = -rel;
        flip = true;
    }
    const int want = pr.sub ? -1 : 1;
    std::vector<Row> fresh(st.n_out);
    for (int o = 0; o < st.n_out; ++o) {
        Row &ra = st.expr[a][o];
        Row &rb = st.expr[b][o];
        for (size_t ia = 0; ia < ra.size(); ++ia) {
            if (ra[ia] == 0)
                continue;
            int sa = dshift(ra[ia]), ga = dsign(ra[ia]);
            int sb = sa + rel;
            int ib = -1;
            for (size_t t = 0; t < rb.size(); ++t)
                if (dshift(rb[t]) == sb) { // a removed digit is 0 -> dshift -1, never matches
                    ib = (int)t;
                    break;
                }
            if (sb >= st.n_bits || ib < 0)
                continue;
            int gb = dsign(rb[ib]);
            if (want * gb * ga != 1)
                continue;
            fresh[o].push_back(flip ? (int8_t)(gb * (sb + 1)) : (int8_t)(ga * (sa + 1)));
            ra[ia] = 0;
            rb[ib] = 0;
        }
        ra.erase(std::remove(ra.begin(), ra.end(), (int8_t)0), ra.end());
        if (a != b)
            rb.erase(std::remove(rb.begin(), rb.end(), (int8_t)0), rb.end());
    }
    st.expr.push_back(std::move(fresh));
}

// state_opr.cc:285-345
static int64_t recount(State &st, const Key &pr) {
    const int64_t id0 = pr.id0, id1 = pr.id1;
    for (auto it = st.freq.begin(); it != st.freq.end();) {
        const Key &k = it->first;
        if (k.id0 == id0 || k.id0 == id1 || k.id1 == id0 || k.id1 == id1)
            it = st.freq.erase(it);
        else
            ++it;
    }
    const int64_t n_c = (int64_t)st.expr.size();
    std::vector<int64_t> modified = {n_c - 1, id0};
    if (id0 != id1)
        modified.push_back(id1);
    std::vector<Key> raw;
    for (int o = 0; o < st.n_out; ++o)
        for (int64_t x = 0; x < n_c; ++x)
            for (int64_t m : modified) {
                if ((x == n_c - 1 || x == id0 || x == id1) && m > x)
                    continue;
                pairs_of(st, std::min(m, x), std::max(m, x), o, raw);
            }
    return add_counted(st, raw);
}

// cmvm_core.cc:10-72
static State greedy(const Mat &kernel, int n_in, int n_out, const std::string &method, const std::vector<QI> &q, const std::vector<float> &lat, int adder_size, int carry_size, Stage *cnt, int64_t max_iters = -1, double time_limit = 0.0) {
    State st = create_state(kernel, n_in, n_out, q, lat, false, cnt);
    auto t0 = std::chrono::steady_clock::now(); // the time budget applies to the greedy loop only
    if (method != "mc" && method != "mc-dc" && method != "mc-pdc" && method != "wmc" && method != "wmc-dc" && method != "wmc-pdc" && method != "dummy") {
        if (!st.freq.empty())
            throw std::runtime_error("Unknown method: " + method);
        return st;
    }
    int64_t it = 0;
    while (!st.freq.empty() && method != "dummy") {
        if (max_iters >= 0 && it >= max_iters)
            break;
        if (time_limit > 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > time_limit)
            break;
        Key k;
        int64_t fsz = (int64_t)st.freq.size();
        if (!select_pair(st, method, k))
            break;
        substitute(st, k, adder_size, carry_size);
        int64_t r = recount(st, k);
        if (cnt) {
            cnt->T += 1;
            cnt->sum_f += fsz;
            cnt->sum_r += r;
        }
        ++it;
    }
    return st;
}

// cmvm_core.cc:75-225
struct HeapEnt {
    float lat;
    int64_t sub, la;
    float qmin, qmax, qstep;
    int64_t id, shift;
    auto tup() const { return std::tie(lat, sub, la, qmin, qmax, qstep, id, shift); }
    bool operator>(const HeapEnt &o) const { return tup() > o.tup(); }
};
static int64_t left_align(const QI &q, int64_t shift) {
    return static_cast<int64_t>(std::log2(std::max(std::abs(q.max + q.step), std::abs(q.min)))) + shift;
}
static Stage finish(const State &st, int adder_size, int carry_size) {
    Stage out;
    out.n_in = st.n_in;
    out.n_out = st.n_out;
    out.carry_size = carry_size;
    out.adder_size = adder_size;
    out.ops = st.ops;
    out.inp_shifts.assign(st.shift0.begin(), st.shift0.end());
    int64_t gid = (int64_t)out.ops.size();
    for (int o = 0; o < st.n_out; ++o) {
        std::priority_queue<HeapEnt, std::vector<HeapEnt>, std::greater<HeapEnt>> heap;
        for (size_t e = 0; e < st.expr.size(); ++e)
            for (int8_t v : st.expr[e][o]) {
                const QI &q = out.ops[e].q;
                int64_t sh = dshift(v);
                heap.push(HeapEnt{out.ops[e].lat, dsign(v) == -1 ? 1 : 0, left_align(q, sh), q.min, q.max, q.step, (int64_t)e, sh});
            }
        if (heap.empty()) {
            out.out_idxs.push_back(-1);
            out.out_shifts.push_back(st.shift1[o]);
            out.out_negs.push_back(0);
            continue;
        }
        if (heap.size() == 1) {
            HeapEnt e = heap.top();
            out.out_idxs.push_back(e.id);
            out.out_shifts.push_back(st.shift1[o] + e.shift);
            out.out_negs.push_back(e.sub);
            continue;
        }
        while (heap.size() > 1) {
            HeapEnt e0 = heap.top();
            heap.pop();
            HeapEnt e1 = heap.top();
            heap.pop();
            QI q0{e0.qmin, e0.qmax, e0.qstep}, q1{e1.qmin, e1.qmax, e1.qstep};
            OpRec op;
            int64_t rshift;
            if (e0.sub) {
                int64_t s = e0.shift - e1.shift;
                QI q = qint_add(q1, q0, s, e1.sub != 0, e0.sub != 0);
                auto [dl, dc] = cost_add(q1, q0, s, (1 ^ e1.sub) != 0, adder_size, carry_size);
                op = OpRec{e1.id, e0.id, 1 ^ e1.sub, s, q, std::max(e0.lat, e1.lat) + dl, dc};
                rshift = e1.shift;
            }
            else {
                int64_t s = e1.shift - e0.shift;
                QI q = qint_add(q0, q1, s, e0.sub != 0, e1.sub != 0);
                auto [dl, dc] = cost_add(q0, q1, s, e1.sub != 0, adder_size, carry_size);
                op = OpRec{e0.id, e1.id, e1.sub, s, q, std::max(e0.lat, e1.lat) + dl, dc};
                rshift = e0.shift;
            }
            heap.push(HeapEnt{op.lat, e0.sub & e1.sub, left_align(op.q, rshift), op.q.min, op.q.max, op.q.step, gid, rshift});
            out.ops.push_back(op);
            ++gid;
        }
        HeapEnt fin = heap.top();
        out.out_idxs.push_back(gid - 1);
        out.out_negs.push_back(fin.sub);
        out.out_shifts.push_back(st.shift1[o] + fin.shift);
    }
    for (auto &e : st.expr)
        for (auto &r : e)
            out.d_final += (int64_t)r.size();
    return out;
}

// cmvm_core.cc:227-237
static Stage solve_single(const Mat &k, int n_in, int n_out, const std::string &method, const std::vector<QI> &q, const std::vector<float> &lat, int adder_size, int carry_size) {
    Stage cnt;
    State st = greedy(k, n_in, n_out, method, q, lat, adder_size, carry_size, &cnt);
    Stage out = finish(st, adder_size, carry_size);
    out.T = cnt.T, out.sum_f = cnt.sum_f, out.sum_r = cnt.sum_r, out.f0 = cnt.f0, out.r0 = cnt.r0, out.d0 = cnt.d0;
    return out;
}

// ---------------------------------------------------------------- mat_decompose.cc:6-137
static std::vector<std::pair<int, int>> prim_mst(const std::vector<int64_t> &cost, int N, int dc) {
    std::vector<float> latm((size_t)N * N);
    for (size_t e = 0; e < latm.size(); ++e)
        latm[e] = std::ceil(std::log2((float)std::max<int64_t>(cost[e], 1)));
    std::vector<int> parent(N, -2), latency(N, 0);
    parent[0] = -1;
    float _dc = -1;
    if (dc >= 0) {
        int64_t mc = cost[0];
        for (int j = 0; j < N; ++j)
            mc = std::max(mc, cost[j]);
        float max_cost0 = (float)mc;
        _dc = (std::pow(2.0, dc) - 1) + std::ceil(std::log2(max_cost0 + 1e-32));
    }
    std::vector<std::pair<int, int>> mapping;
    for (int n_impl = 1; n_impl < N; ++n_impl) {
        int64_t best = std::numeric_limits<int64_t>::max();
        int bi = -1, bj = -1;
        for (int i = 0; i < N; ++i) {
            if (parent[i] != -2)
                continue;
            for (int j = 0; j < N; ++j) {
                if (parent[j] == -2)
                    continue;
                int64_t c = cost[(size_t)i * N + j];
                if (dc >= 0) {
                    float ml = std::max(latm[(size_t)i * N + j], (float)latency[j]) + 1;
                    if (ml > _dc)
                        c = std::numeric_limits<int64_t>::max() / 2;
                }
                if (c < best) {
                    best = c;
                    bi = i;
                    bj = j;
                }
            }
        }
        parent[bi] = bj;
        mapping.push_back({bj, bi});
        latency[bi] = (int)(std::max(latm[(size_t)bi * N + bj], (float)latency[bj]) + 1);
    }
    return mapping;
}

static void kernel_decompose(const Mat &kernel, int n_in, int n_out, int dc, Mat &m0, Mat &m1) {
    Mat c(kernel);
    std::vector<int8_t> s0, s1;
    center(c, n_in, n_out, s0, s1);
    const int n = n_out + 1;
    Mat aug((size_t)n_in * n, 0.0f);
    for (int i = 0; i < n_in; ++i)
        for (int j = 0; j < n_out; ++j)
            aug[(size_t)i * n + j + 1] = c[(size_t)i * n_out + j];
    // CSD weight of every pairwise difference / sum (each tensor with its own width N)
    std::vector<int32_t> d0((size_t)n_in * n * n), d1((size_t)n_in * n * n);
    for (int i = 0; i < n_in; ++i)
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) {
                d0[((size_t)i * n + a) * n + b] = (int32_t)(aug[(size_t)i * n + a] - aug[(size_t)i * n + b]);
                d1[((size_t)i * n + a) * n + b] = (int32_t)(aug[(size_t)i * n + a] + aug[(size_t)i * n + b]);
            }
    std::vector<int64_t> dist0((size_t)n * n, 0), dist1((size_t)n * n, 0), dist((size_t)n * n), sign((size_t)n * n);
    for (int which = 0; which < 2; ++which) {
        std::vector<int8_t> dg;
        int N = to_csd(which ? d1 : d0, dg);
        auto &dst = which ? dist1 : dist0;
        for (int i = 0; i < n_in; ++i)
            for (int ab = 0; ab < n * n; ++ab)
                for (int k = 0; k < N; ++k)
                    dst[ab] += dg[(((size_t)i * n * n) + ab) * N + k] != 0;
    }
    for (size_t e = 0; e < dist.size(); ++e) {
        sign[e] = (dist1[e] - dist0[e] < 0) ? -1 : 1;
        dist[e] = std::min(dist0[e], dist1[e]);
    }
    auto mapping = prim_mst(dist, n, dc);
    m0.assign((size_t)n_in * n_out, 0.0f);
    m1.assign((size_t)n_out * n_out, 0.0f);
    std::vector<float> sc0(n_in), sc1(n_out);
    for (int i = 0; i < n_in; ++i)
        sc0[i] = std::pow(2.0f, (float)s0[i]);
    for (int j = 0; j < n_out; ++j)
        sc1[j] = std::pow(2.0f, (float)s1[j]);
    if (dc == -1) {
        for (int i = 0; i < n_in; ++i)
            for (int j = 0; j < n_out; ++j)
                m0[(size_t)i * n_out + j] = c[(size_t)i * n_out + j] * sc0[i];
        for (int j = 0; j < n_out; ++j)
            m1[(size_t)j * n_out + j] = 1.0f * sc1[j];
        return;
    }
    int cnt = 0;
    for (auto [from, to] : mapping) {
        float sg = (float)sign[(size_t)to * n + from];
        std::vector<float> col0(n_in), col1(n_out, 0.0f);
        bool any = false;
        for (int i = 0; i < n_in; ++i) {
            col0[i] = aug[(size_t)i * n + to] - aug[(size_t)i * n + from] * sg;
            any |= col0[i] != 0.0f;
        }
        if (from != 0)
            for (int j = 0; j < n_out; ++j)
                col1[j] = m1[(size_t)j * n_out + from - 1] * sg;
        if (any) {
            col1[cnt] = 1.0f;
            for (int i = 0; i < n_in; ++i)
                m0[(size_t)i * n_out + cnt] = col0[i];
            ++cnt;
        }
        for (int j = 0; j < n_out; ++j)
            m1[(size_t)j * n_out + to - 1] = col1[j];
    }
    for (int i = 0; i < n_in; ++i)
        for (int j = 0; j < n_out; ++j)
            m0[(size_t)i * n_out + j] *= sc0[i];
    for (int i = 0; i < n_out; ++i)
        for (int j = 0; j < n_out; ++j)
            m1[(size_t)i * n_out + j] *= sc1[j];
}

// ---------------------------------------------------------------- api.cc:11-250
static bool ends_with(const std::string &s, const std::string &t) { return s.size() >= t.size() && s.compare(s.size() - t.size(), t.size(), t) == 0; }
static float stage_max_lat(const Stage &s) {
    float m = 0.0f;
    for (int64_t idx : s.out_idxs)
        m = std::max(m, idx >= 0 ? s.ops[idx].lat : 0.0f);
    return m;
}
static float minimal_latency(const Mat &k, int n_in, int n_out, const std::vector<QI> &q, const std::vector<float> &lat, int carry_size, int adder_size) {
    State st = create_state(k, n_in, n_out, q, lat, true, nullptr);
    return stage_max_lat(finish(st, adder_size, carry_size));
}
static std::vector<Stage> solve_one(const Mat &kernel, int n_in, int n_out, std::string method0, std::string method1, int hard_dc, int decompose_dc, const std::vector<QI> &q, const std::vector<float> &lat, int adder_size, int carry_size) {
    if (method1 == "auto")
        method1 = (hard_dc >= 6 || ends_with(method0, "dc")) ? method0 : method0 + "-dc";
    if (hard_dc == 0 && !ends_with(method0, "dc"))
        method0 = method0 + "-dc";
    float min_lat = std::numeric_limits<float>::infinity();
    if (hard_dc >= 0)
        min_lat = minimal_latency(kernel, n_in, n_out, q, lat, carry_size, adder_size);
    float allowed = hard_dc + min_lat;
    int log2_n = (int)std::ceil(std::log2((float)n_in));
    decompose_dc = decompose_dc == -2 ? std::min(hard_dc, log2_n) : std::min({hard_dc, decompose_dc, log2_n});
    Stage s0, s1;
    while (true) {
        if (decompose_dc < 0 && hard_dc >= 0) {
            if (method0 != "dummy")
                method0 = method1 = "wmc-dc";
            else
                method0 = method1 = "dummy";
        }
        Mat m0, m1;
        kernel_decompose(kernel, n_in, n_out, decompose_dc, m0, m1);
        s0 = solve_single(m0, n_in, n_out, method0, q, lat, adder_size, carry_size);
        std::vector<QI> q0;
        std::vector<float> l0;
        for (int64_t idx : s0.out_idxs) {
            l0.push_back(idx >= 0 ? s0.ops[idx].lat : 0.0f);
            q0.push_back(idx >= 0 ? s0.ops[idx].q : QI{0.0f, 0.0f, std::numeric_limits<float>::infinity()});
        }
        const bool both = method0 == "wmc-dc" && method1 == "wmc-dc";
        if (stage_max_lat(s0) > allowed && (!both || decompose_dc >= 0)) {
            decompose_dc--;
            continue;
        }
        s1 = solve_single(m1, n_out, n_out, method1, q0, l0, adder_size, carry_size);
        if (stage_max_lat(s1) > allowed && (!both || decompose_dc >= 0)) {
            decompose_dc--;
            continue;
        }
        break;
    }
    return {s0, s1};
}
static std::vector<Stage> solve(const Mat &kernel, int n_in, int n_out, const std::string &method0, const std::string &method1, int hard_dc, int decompose_dc, std::vector<QI> q, std::vector<float> lat, int adder_size, int carry_size, bool search_all) {
    if (q.empty())
        q.assign(n_in, QI{-128.0f, 127.0f, 1.0f});
    if (lat.empty())
        lat.assign(n_in, 0.0f);
    if (!search_all)
        return solve_one(kernel, n_in, n_out, method0, method1, hard_dc, decompose_dc, q, lat, adder_size, carry_size);
    int hd = hard_dc < 0 ? 1000000000 : hard_dc;
    int max_dc = std::min(hd, (int)std::ceil(std::log2((float)n_in)));
    std::vector<int> tries;
    for (int d = -1; d <= max_dc; ++d)
        tries.push_back(d);
    std::vector<std::vector<Stage>> cand(tries.size());
    std::vector<float> costs(tries.size());
    std::string err;
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < tries.size(); ++i) {
        try {
            cand[i] = solve_one(kernel, n_in, n_out, method0, method1, hd, tries[i], q, lat, adder_size, carry_size);
            float c = 0.0f;
            for (auto &s : cand[i])
                for (auto &op : s.ops)
                    c += op.cost;
            costs[i] = c;
        }
        catch (const std::exception &e) {
#pragma omp critical
            err = e.what();
        }
    }
    if (!err.empty())
        throw std::runtime_error(err);
    size_t best = 0;
    for (size_t i = 1; i < tries.size(); ++i)
        if (costs[i] < costs[best])
            best = i;
    return cand[best];
}

thread_local std::string g_err;
struct Handle {
    std::vector<Stage> stages;
};
static Mat to_mat(const float *k, int64_t n_in, int64_t n_out) { return Mat(k, k + n_in * n_out); }
static std::vector<QI> to_q(const float *q, int64_t n) {
    std::vector<QI> v;
    if (q)
        for (int64_t i = 0; i < n; ++i)
            v.push_back(QI{q[3 * i], q[3 * i + 1], q[3 * i + 2]});
    return v;
}

} // namespace orc

using namespace orc;
extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

void *orc_solve(const float *kernel, int64_t n_in, int64_t n_out, const char *method0, const char *method1, int hard_dc, int decompose_dc, const float *qint, const float *lat, int adder_size, int carry_size, int search_all) {
    try {
        auto *h = new Handle();
        h->stages = solve(to_mat(kernel, n_in, n_out), (int)n_in, (int)n_out, method0, method1, hard_dc, decompose_dc, to_q(qint, n_in), lat ? std::vector<float>(lat, lat + n_in) : std::vector<float>{}, adder_size, carry_size, search_all != 0);
        return h;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void *orc_solve_single(const float *kernel, int64_t n_in, int64_t n_out, const char *method, const float *qint, const float *lat, int adder_size, int carry_size) {
    try {
        auto q = to_q(qint, n_in);
        if (q.empty())
            q.assign(n_in, QI{-128.0f, 127.0f, 1.0f});
        std::vector<float> l = lat ? std::vector<float>(lat, lat + n_in) : std::vector<float>(n_in, 0.0f);
        auto *h = new Handle();
        h->stages.push_back(solve_single(to_mat(kernel, n_in, n_out), (int)n_in, (int)n_out, method, q, l, adder_size, carry_size));
        return h;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
// Bounded run of the greedy loop for CPU-baseline sampling: out = {iterations, seconds, sum_f, sum_r, f0, r0, d0}
int orc_partial(const float *kernel, int64_t n_in, int64_t n_out, const char *method, int64_t max_iters, double time_limit, double *out) {
    try {
        std::vector<QI> q(n_in, QI{-128.0f, 127.0f, 1.0f});
        std::vector<float> l(n_in, 0.0f);
        Stage cnt;
        auto t0 = std::chrono::steady_clock::now();
        greedy(to_mat(kernel, n_in, n_out), (int)n_in, (int)n_out, method, q, l, -1, -1, &cnt, max_iters, time_limit);
        out[0] = (double)cnt.T;
        out[1] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out[2] = (double)cnt.sum_f;
        out[3] = (double)cnt.sum_r;
        out[4] = (double)cnt.f0;
        out[5] = (double)cnt.r0;
        out[6] = (double)cnt.d0;
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
void orc_free(void *h) { delete static_cast<Handle *>(h); }
int64_t orc_n_stages(void *h) { return (int64_t) static_cast<Handle *>(h)->stages.size(); }
int64_t orc_stage_n_ops(void *h, int64_t s) { return (int64_t) static_cast<Handle *>(h)->stages[s].ops.size(); }
void orc_stage_meta(void *h, int64_t s, int64_t *out) {
    auto &c = static_cast<Handle *>(h)->stages[s];
    out[0] = c.n_in, out[1] = c.n_out, out[2] = c.carry_size, out[3] = c.adder_size;
    out[4] = c.T, out[5] = c.sum_f, out[6] = c.sum_r, out[7] = c.f0, out[8] = c.r0, out[9] = c.d0, out[10] = c.d_final;
}
void orc_stage_copy(void *h, int64_t s, int64_t *inp_shifts, int64_t *out_idxs, int64_t *out_shifts, int64_t *out_negs, int64_t *ops_i, float *ops_f) {
    auto &c = static_cast<Handle *>(h)->stages[s];
    std::copy(c.inp_shifts.begin(), c.inp_shifts.end(), inp_shifts);
    std::copy(c.out_idxs.begin(), c.out_idxs.end(), out_idxs);
    std::copy(c.out_shifts.begin(), c.out_shifts.end(), out_shifts);
    std::copy(c.out_negs.begin(), c.out_negs.end(), out_negs);
    for (size_t i = 0; i < c.ops.size(); ++i) {
        const OpRec &op = c.ops[i];
        ops_i[4 * i] = op.id0, ops_i[4 * i + 1] = op.id1, ops_i[4 * i + 2] = op.opcode, ops_i[4 * i + 3] = op.data;
        ops_f[5 * i] = op.q.min, ops_f[5 * i + 1] = op.q.max, ops_f[5 * i + 2] = op.q.step, ops_f[5 * i + 3] = op.lat, ops_f[5 * i + 4] = op.cost;
    }
}
int64_t orc_csd_decompose(const float *kernel, int64_t n_in, int64_t n_out, int do_center, int8_t *csd, int8_t *shift0, int8_t *shift1) {
    std::vector<int8_t> d, s0, s1;
    int N = csd_decompose(to_mat(kernel, n_in, n_out), (int)n_in, (int)n_out, do_center != 0, d, s0, s1);
    std::copy(d.begin(), d.end(), csd);
    std::copy(s0.begin(), s0.end(), shift0);
    std::copy(s1.begin(), s1.end(), shift1);
    return N;
}
int64_t orc_int_arr_to_csd(const int32_t *x, int64_t n, int8_t *out) {
    std::vector<int8_t> d;
    int N = to_csd(std::vector<int32_t>(x, x + n), d);
    std::copy(d.begin(), d.end(), out);
    return N;
}
void orc_kernel_decompose(const float *kernel, int64_t n_in, int64_t n_out, int dc, float *m0, float *m1) {
    Mat a, b;
    kernel_decompose(to_mat(kernel, n_in, n_out), (int)n_in, (int)n_out, dc, a, b);
    std::copy(a.begin(), a.end(), m0);
    std::copy(b.begin(), b.end(), m1);
}
int orc_get_lsb_loc(float x) { return lsb_loc(x); }
int orc_iceil_log2(float x) { return iceil_log2(x); }
void orc_cost_add(const float *q0, const float *q1, int64_t shift, int sub, int adder_size, int carry_size, float *out) {
    auto [l, c] = cost_add(QI{q0[0], q0[1], q0[2]}, QI{q1[0], q1[1], q1[2]}, shift, sub != 0, adder_size, carry_size);
    out[0] = l, out[1] = c;
}
float orc_log2f(float x) { return std::log2(x); }
int orc_csd_weight(int32_t x) {
    std::vector<int8_t> d;
    int N = to_csd(std::vector<int32_t>{x}, d);
    int w = 0;
    for (int k = 0; k < N; ++k)
        w += d[k] != 0;
    return w;
}
int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} // extern "C"
