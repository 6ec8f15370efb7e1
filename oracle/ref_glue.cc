// oracle/ref_glue.cc — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Builds, together with the reference's own *unmodified* translation units compiled in place
// from /root/reference/src/da4ml/_binary/cmvm/{api,cmvm_core,state_opr,indexers}.cc, the shared
// library oracle/_ref/libcmvm_ref.so (see oracle/Makefile).
//
// Two reference files cannot be compiled here because they are written as xtensor expression
// templates and xtensor/xsimd/xtl are network-fetched meson wraps that are absent from this
// image (subprojects/xtensor.wrap): bit_decompose.cc and mat_decompose.cc.  The functions those
// files define -- and that the four compiled TUs link against -- are restated below as plain
// loops over the stand-in xt::xarray (oracle/shim), each citing the reference lines it follows.
// Everything else (the greedy CSE loop, selectors, histogram, to_solution, _solve/solve) is the
// reference's own object code.
//
// The extern "C" block at the bottom is a ctypes-friendly window onto the reference API, plus
// step-level tracing built only from the reference's public functions
// (create_state / idx_* / update_state in state_opr.hh, indexers.hh).

#include "api.hh"
#include "cmvm_core.hh"
#include "indexers.hh"
#include "mat_decompose.hh"
#include "state_opr.hh"

#include <bit>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <tuple>

// ---------------------------------------------------------------------------------------------
// bit_decompose.cc:10-20
int8_t get_lsb_loc(float x) {
    if (x == 0.0f)
        return 127;
    uint32_t bits = std::bit_cast<uint32_t>(x);
    uint8_t exp = static_cast<uint8_t>((bits >> 23) & 0xFF);
    uint32_t mant = bits & 0x7FFFFF;
    int mtz = __builtin_ctz(mant + (1 << 23));
    return static_cast<int8_t>(exp + mtz - 150);
}

// bit_decompose.cc:22-42.  `x` is consumed (the reference reshapes and mutates it in place).
xt::xarray<int8_t> _volatile_int_arr_to_csd(xt::xarray<int32_t> &x) {
    int32_t max_val = 0;
    for (int32_t v : x)
        max_val = std::max(max_val, static_cast<int32_t>(std::abs(v)));
    size_t N = static_cast<size_t>(
        std::ceil(std::log2(std::max(static_cast<float>(max_val), 1.0f) * 1.5))
    );
    N = std::max(N, size_t(1));
    auto out_shape = x.shape();
    out_shape.push_back(N);
    xt::xarray<int8_t> buf(out_shape, 0);
    int32_t *xs = x.data();
    int8_t *out = buf.data();
    size_t n_el = x.size();
    for (int n = static_cast<int>(N) - 1; n >= 0; --n) {
        int32_t _2pn = static_cast<int32_t>(1U << n);
        int32_t thres = _2pn * 2 / 3;
        for (size_t e = 0; e < n_el; ++e) {
            int8_t d = static_cast<int8_t>(xs[e] > thres) - static_cast<int8_t>(xs[e] < -thres);
            out[e * N + n] = d;
            xs[e] -= _2pn * static_cast<int32_t>(d);
        }
    }
    return buf;
}

// bit_decompose.hh:21-34 (_shift_amount, _center).  Column shifts first, then row shifts on the
// already column-scaled array.  2-D only.
static std::tuple<xt::xarray<float>, xt::xarray<int8_t>, xt::xarray<int8_t>>
center_restated(xt::xarray<float> &arr) {
    if (arr.dimension() != 2)
        throw std::runtime_error("csd_decompose only supports 2D arrays.");
    size_t n_in = arr.shape(0), n_out = arr.shape(1);
    xt::xarray<int8_t> shift1({n_out}, 0), shift0({n_in}, 0);
    for (size_t j = 0; j < n_out; ++j) {
        int8_t m = 127;
        for (size_t i = 0; i < n_in; ++i)
            m = std::min(m, get_lsb_loc(arr(i, j)));
        shift1(j) = m;
    }
    for (size_t i = 0; i < n_in; ++i)
        for (size_t j = 0; j < n_out; ++j)
            arr(i, j) = static_cast<float>(arr(i, j) * std::pow(2.0, -static_cast<int>(shift1(j))));
    for (size_t i = 0; i < n_in; ++i) {
        int8_t m = 127;
        for (size_t j = 0; j < n_out; ++j)
            m = std::min(m, get_lsb_loc(arr(i, j)));
        shift0(i) = m;
    }
    for (size_t i = 0; i < n_in; ++i)
        for (size_t j = 0; j < n_out; ++j)
            arr(i, j) = static_cast<float>(arr(i, j) * std::pow(2.0, -static_cast<int>(shift0(i))));
    return {arr, shift0, shift1};
}

// bit_decompose.cc:44-62
std::tuple<xt::xarray<int8_t>, xt::xarray<int8_t>, xt::xarray<int8_t>>
csd_decompose(xt::xarray<float> &arr, bool center) {
    xt::xarray<float> arr_cpy(arr);
    if (arr_cpy.dimension() != 2)
        throw std::runtime_error("csd_decompose only supports 2D arrays.");
    xt::xarray<int8_t> shift0({arr_cpy.shape(0)}, 0);
    xt::xarray<int8_t> shift1({arr_cpy.shape(1)}, 0);
    if (center)
        std::tie(arr_cpy, shift0, shift1) = center_restated(arr_cpy);
    xt::xarray<int32_t> arr_int(arr_cpy.shape(), 0);
    for (size_t e = 0; e < arr_cpy.size(); ++e)
        arr_int.data()[e] = static_cast<int32_t>(arr_cpy.data()[e]);
    auto csd = _volatile_int_arr_to_csd(arr_int);
    return {csd, shift0, shift1};
}

// mat_decompose.cc:6-60
xt::xarray<int32_t> prim_mst_dc(const xt::xarray<int64_t> &cost_mat, int dc) {
    size_t N = cost_mat.shape(0);
    xt::xarray<float> lat_mat({N, N}, 0.0f);
    for (size_t i = 0; i < N; ++i)
        for (size_t j = 0; j < N; ++j)
            lat_mat(i, j) = std::ceil(
                std::log2(static_cast<float>(std::max<int64_t>(cost_mat(i, j), 1)))
            );
    std::vector<int32_t> parent(N, -2);
    parent[0] = -1;
    xt::xarray<int32_t> mapping({N - 1, size_t(2)}, 0);
    std::vector<int32_t> latency(N, 0);

    float _dc = -1;
    if (dc >= 0) {
        int64_t mc = cost_mat(0, 0);
        for (size_t j = 0; j < N; ++j)
            mc = std::max(mc, cost_mat(0, j));
        float max_cost0 = static_cast<float>(mc);
        _dc = (std::pow(2.0, dc) - 1) + std::ceil(std::log2(max_cost0 + 1e-32));
    }

    for (size_t n_impl = 1; n_impl < N; ++n_impl) {
        std::vector<size_t> not_impl, impl;
        for (size_t i = 0; i < N; ++i) {
            if (parent[i] != -2)
                impl.push_back(i);
            else
                not_impl.push_back(i);
        }
        int64_t best_cost = std::numeric_limits<int64_t>::max();
        size_t best_i = 0, best_j = 0;
        for (size_t ii = 0; ii < not_impl.size(); ++ii) {
            for (size_t jj = 0; jj < impl.size(); ++jj) {
                size_t i = not_impl[ii], j = impl[jj];
                int64_t c = cost_mat(i, j);
                if (dc >= 0) {
                    float lat = lat_mat(i, j);
                    float max_lat = std::max(lat, (float)latency[j]) + 1;
                    if (max_lat > _dc)
                        c = std::numeric_limits<int64_t>::max() / 2;
                }
                if (c < best_cost) {
                    best_cost = c;
                    best_i = ii;
                    best_j = jj;
                }
            }
        }
        size_t i = not_impl[best_i], j = impl[best_j];
        parent[i] = static_cast<int32_t>(j);
        mapping(n_impl - 1, 0) = static_cast<int32_t>(j);
        mapping(n_impl - 1, 1) = static_cast<int32_t>(i);
        latency[i] = static_cast<int32_t>(std::max(lat_mat(i, j), (float)latency[j]) + 1);
    }
    return mapping;
}

// mat_decompose.cc:62-137
std::pair<xt::xarray<float>, xt::xarray<float>> kernel_decompose(xt::xarray<float> kernel, int dc) {
    auto [centered, shift0, shift1] = center_restated(kernel);
    size_t n_in = centered.shape(0), n_out = centered.shape(1);
    std::vector<float> scale0(n_in), scale1(n_out);
    for (size_t i = 0; i < n_in; ++i)
        scale0[i] = std::pow(2.0f, static_cast<float>(shift0(i)));
    for (size_t j = 0; j < n_out; ++j)
        scale1[j] = std::pow(2.0f, static_cast<float>(shift1(j)));

    size_t m = n_in, n = n_out + 1;
    xt::xarray<float> mat_aug({m, n}, 0.0f);
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n_out; ++j)
            mat_aug(i, j + 1) = centered(i, j);

    xt::xarray<int32_t> diff0_int({m, n, n}, 0), diff1_int({m, n, n}, 0);
    for (size_t i = 0; i < m; ++i)
        for (size_t a = 0; a < n; ++a)
            for (size_t b = 0; b < n; ++b) {
                diff0_int(i, a, b) = static_cast<int32_t>(mat_aug(i, a) - mat_aug(i, b));
                diff1_int(i, a, b) = static_cast<int32_t>(mat_aug(i, a) + mat_aug(i, b));
            }
    xt::xarray<int64_t> dist0({n, n}, 0), dist1({n, n}, 0);
    {
        auto csd0 = _volatile_int_arr_to_csd(diff0_int);
        size_t N0 = csd0.shape(3);
        const int8_t *p = csd0.data();
        for (size_t i = 0; i < m; ++i)
            for (size_t a = 0; a < n; ++a)
                for (size_t b = 0; b < n; ++b)
                    for (size_t k = 0; k < N0; ++k)
                        dist0(a, b) += (*p++ != 0);
    }
    {
        auto csd1 = _volatile_int_arr_to_csd(diff1_int);
        size_t N1 = csd1.shape(3);
        const int8_t *p = csd1.data();
        for (size_t i = 0; i < m; ++i)
            for (size_t a = 0; a < n; ++a)
                for (size_t b = 0; b < n; ++b)
                    for (size_t k = 0; k < N1; ++k)
                        dist1(a, b) += (*p++ != 0);
    }
    xt::xarray<int64_t> sign_arr({n, n}, 1), dist({n, n}, 0);
    for (size_t a = 0; a < n; ++a)
        for (size_t b = 0; b < n; ++b) {
            sign_arr(a, b) = (dist1(a, b) - dist0(a, b) < 0) ? -1 : 1;
            dist(a, b) = std::min(dist0(a, b), dist1(a, b));
        }

    auto mapping_arr = prim_mst_dc(dist, dc);

    xt::xarray<float> m0({n_in, n_out}, 0.0f), m1({n_out, n_out}, 0.0f);
    if (dc == -1) {
        for (size_t i = 0; i < n_in; ++i)
            for (size_t j = 0; j < n_out; ++j)
                m0(i, j) = centered(i, j) * scale0[i];
        for (size_t j = 0; j < n_out; ++j)
            m1(j, j) = 1.0f * scale1[j];
        return {m0, m1};
    }

    size_t cnt = 0;
    for (size_t k = 0; k < mapping_arr.shape(0); ++k) {
        int32_t _from = mapping_arr(k, 0);
        int32_t _to = mapping_arr(k, 1);
        float sgn = static_cast<float>(sign_arr(_to, _from));
        std::vector<float> col0(n_in), col1(n_out, 0.0f);
        bool any = false;
        for (size_t i = 0; i < n_in; ++i) {
            col0[i] = mat_aug(i, _to) - mat_aug(i, _from) * sgn;
            any = any || (col0[i] != 0.0f);
        }
        if (_from != 0)
            for (size_t j = 0; j < n_out; ++j)
                col1[j] = m1(j, _from - 1) * sgn;
        if (any) {
            col1[cnt] = 1.0f;
            for (size_t i = 0; i < n_in; ++i)
                m0(i, cnt) = col0[i];
            cnt++;
        }
        for (size_t j = 0; j < n_out; ++j)
            m1(j, _to - 1) = col1[j];
    }
    for (size_t i = 0; i < n_in; ++i)
        for (size_t j = 0; j < n_out; ++j)
            m0(i, j) = m0(i, j) * scale0[i];
    for (size_t i = 0; i < n_out; ++i)
        for (size_t j = 0; j < n_out; ++j)
            m1(i, j) = m1(i, j) * scale1[j];
    return {m0, m1};
}

// ---------------------------------------------------------------------------------------------
// ctypes window
namespace {

thread_local std::string g_err;

xt::xarray<float> to_xarr(const float *k, int64_t n_in, int64_t n_out) {
    xt::xarray<float> a({(size_t)n_in, (size_t)n_out}, 0.0f);
    std::memcpy(a.data(), k, sizeof(float) * n_in * n_out);
    return a;
}
std::vector<QInterval> to_qints(const float *q, int64_t n) {
    std::vector<QInterval> v;
    if (q)
        for (int64_t i = 0; i < n; ++i)
            v.push_back(QInterval{q[3 * i], q[3 * i + 1], q[3 * i + 2]});
    return v;
}
std::vector<float> to_lats(const float *l, int64_t n) {
    return l ? std::vector<float>(l, l + n) : std::vector<float>{};
}

struct RefTrace {
    std::vector<int64_t> pairs; // 4 per step: id0,id1,shift,sub
    std::vector<int64_t> f_sizes; // |F_t| scanned by the selector at step t
    std::vector<int64_t> r_sizes; // raw pairs enumerated by update_stats at step t
    int64_t f0 = 0, r0 = 0, d0 = 0, d_final = 0;
    double seconds = 0;
    double create_seconds = 0;
};

// number of digit pairs update_stats (state_opr.cc:307-340) enumerates for `pair` on the
// post-update_expr state
int64_t count_raw_pairs(const DAState &st, const Pair &pair) {
    int64_t id0 = pair.id0, id1 = pair.id1;
    int64_t n_c = (int64_t)st.expr.size();
    std::vector<int64_t> modified = {n_c - 1, id0};
    if (id0 != id1)
        modified.push_back(id1);
    size_t n_out = st.kernel.shape(1);
    int64_t r = 0;
    for (size_t o = 0; o < n_out; ++o)
        for (int64_t in1 = 0; in1 < n_c; ++in1)
            for (auto in0 : modified) {
                if ((in1 == n_c - 1 || in1 == id0 || in1 == id1) && in0 > in1)
                    continue;
                int64_t lo = std::min(in0, in1), hi = std::max(in0, in1);
                size_t a = st.expr[lo].rows[o].size(), b = st.expr[hi].rows[o].size();
                if (!a || !b)
                    continue;
                r += (lo == hi) ? (int64_t)(a * (a - 1) / 2) : (int64_t)(a * b);
            }
    return r;
}

Pair select(const DAState &state, const std::string &method) {
    if (method == "mc")
        return idx_mc(state);
    if (method == "mc-dc")
        return idx_mc_dc(state, true);
    if (method == "mc-pdc")
        return idx_mc_dc(state, false);
    if (method == "wmc")
        return idx_wmc(state);
    if (method == "wmc-dc")
        return idx_wmc_dc(state, true);
    if (method == "wmc-pdc")
        return idx_wmc_dc(state, false);
    throw std::runtime_error("Unknown method: " + method);
}

} // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// ---- full solve (api.cc:147) -------------------------------------------------------------------
void *ref_solve(
    const float *kernel, int64_t n_in, int64_t n_out, const char *method0, const char *method1,
    int hard_dc, int decompose_dc, const float *qint, const float *lat, int adder_size,
    int carry_size, int search_all
) {
    try {
        auto *res = new PipelineResult(solve(
            to_xarr(kernel, n_in, n_out), method0, method1, hard_dc, decompose_dc,
            to_qints(qint, n_in), to_lats(lat, n_in), adder_size, carry_size, search_all != 0
        ));
        return res;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

// ---- one stage (cmvm_core.cc:227) ---------------------------------------------------------------
void *ref_solve_single(
    const float *kernel, int64_t n_in, int64_t n_out, const char *method, const float *qint,
    const float *lat, int adder_size, int carry_size
) {
    try {
        std::vector<QInterval> q = to_qints(qint, n_in);
        std::vector<float> l = to_lats(lat, n_in);
        if (q.empty())
            q.assign(n_in, QInterval{-128.0, 127.0, 1.0});
        if (l.empty())
            l.assign(n_in, 0.0f);
        auto *res = new PipelineResult();
        res->solutions.push_back(
            solve_single(to_xarr(kernel, n_in, n_out), method, q, l, adder_size, carry_size)
        );
        return res;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void ref_free(void *h) { delete static_cast<PipelineResult *>(h); }
int64_t ref_n_stages(void *h) { return (int64_t) static_cast<PipelineResult *>(h)->solutions.size(); }
int64_t ref_stage_n_ops(void *h, int64_t s) {
    return (int64_t) static_cast<PipelineResult *>(h)->solutions[s].ops.size();
}
void ref_stage_meta(void *h, int64_t s, int64_t *out /*n_in,n_out,carry,adder*/) {
    auto &c = static_cast<PipelineResult *>(h)->solutions[s];
    out[0] = c.shape.first;
    out[1] = c.shape.second;
    out[2] = c.carry_size;
    out[3] = c.adder_size;
}
// ops_i: [n_ops,4] int64 (id0,id1,opcode,data); ops_f: [n_ops,5] float (min,max,step,latency,cost)
void ref_stage_copy(
    void *h, int64_t s, int64_t *inp_shifts, int64_t *out_idxs, int64_t *out_shifts,
    int64_t *out_negs, int64_t *ops_i, float *ops_f
) {
    auto &c = static_cast<PipelineResult *>(h)->solutions[s];
    std::copy(c.inp_shifts.begin(), c.inp_shifts.end(), inp_shifts);
    std::copy(c.out_idxs.begin(), c.out_idxs.end(), out_idxs);
    std::copy(c.out_shifts.begin(), c.out_shifts.end(), out_shifts);
    std::copy(c.out_negs.begin(), c.out_negs.end(), out_negs);
    for (size_t i = 0; i < c.ops.size(); ++i) {
        const Op &op = c.ops[i];
        ops_i[4 * i + 0] = op.id0;
        ops_i[4 * i + 1] = op.id1;
        ops_i[4 * i + 2] = op.opcode;
        ops_i[4 * i + 3] = op.data;
        ops_f[5 * i + 0] = op.qint.min;
        ops_f[5 * i + 1] = op.qint.max;
        ops_f[5 * i + 2] = op.qint.step;
        ops_f[5 * i + 3] = op.latency;
        ops_f[5 * i + 4] = op.cost;
    }
}

// ---- helpers exported by the reference's nanobind module (bindings.cc:227-263) -----------------
int ref_get_lsb_loc(float x) { return get_lsb_loc(x); }
int ref_iceil_log2(float x) { return iceil_log2(x); }
void ref_cost_add(const float *q0, const float *q1, int64_t shift, int sub, int adder_size, int carry_size, float *out) {
    auto [l, c] = cost_add(QInterval{q0[0], q0[1], q0[2]}, QInterval{q1[0], q1[1], q1[2]}, shift, sub != 0, adder_size, carry_size);
    out[0] = l;
    out[1] = c;
}
void ref_qint_add(const float *q0, const float *q1, int64_t shift, int sub0, int sub1, float *out) {
    QInterval q = qint_add(QInterval{q0[0], q0[1], q0[2]}, QInterval{q1[0], q1[1], q1[2]}, shift, sub0 != 0, sub1 != 0);
    out[0] = q.min;
    out[1] = q.max;
    out[2] = q.step;
}
void ref_overlap_and_accum(const float *q0, const float *q1, int *out) {
    auto [a, b] = overlap_and_accum(QInterval{q0[0], q0[1], q0[2]}, QInterval{q1[0], q1[1], q1[2]});
    out[0] = a;
    out[1] = b;
}
// returns N (digits per element); csd must hold n_in*n_out*32 int8 at most; caller reads [.., N]
int64_t ref_csd_decompose(const float *kernel, int64_t n_in, int64_t n_out, int center, int8_t *csd, int8_t *shift0, int8_t *shift1) {
    auto k = to_xarr(kernel, n_in, n_out);
    auto [c, s0, s1] = csd_decompose(k, center != 0);
    std::copy(c.begin(), c.end(), csd);
    std::copy(s0.begin(), s0.end(), shift0);
    std::copy(s1.begin(), s1.end(), shift1);
    return (int64_t)c.shape(2);
}
int64_t ref_int_arr_to_csd(const int32_t *x, int64_t n, int8_t *out) {
    xt::xarray<int32_t> a({(size_t)n}, 0);
    std::copy(x, x + n, a.data());
    auto c = _volatile_int_arr_to_csd(a);
    std::copy(c.begin(), c.end(), out);
    return (int64_t)c.shape(1);
}
void ref_kernel_decompose(const float *kernel, int64_t n_in, int64_t n_out, int dc, float *m0, float *m1) {
    auto [a, b] = kernel_decompose(to_xarr(kernel, n_in, n_out), dc);
    std::copy(a.begin(), a.end(), m0);
    std::copy(b.begin(), b.end(), m1);
}
float ref_log2f(float x) { return std::log2(x); }

// ---- step-level trace of the greedy loop (cmvm_core.cc:10-72 re-driven through the public
//      create_state / idx_* / update_state) -------------------------------------------------------
// Runs at most max_iters iterations (<0: to completion) or until time_limit_s (<=0: none).
void *ref_trace(
    const float *kernel, int64_t n_in, int64_t n_out, const char *method, const float *qint,
    const float *lat, int adder_size, int carry_size, int64_t max_iters, double time_limit_s,
    int want_counters
) {
    try {
        std::vector<QInterval> q = to_qints(qint, n_in);
        std::vector<float> l = to_lats(lat, n_in);
        if (q.empty())
            q.assign(n_in, QInterval{-128.0, 127.0, 1.0});
        if (l.empty())
            l.assign(n_in, 0.0f);
        auto *tr = new RefTrace();
        auto t0 = std::chrono::steady_clock::now();
        DAState state = create_state(to_xarr(kernel, n_in, n_out), q, l);
        tr->create_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        tr->f0 = (int64_t)state.freq_stat.size();
        for (auto &e : state.expr)
            for (auto &r : e.rows)
                tr->d0 += (int64_t)r.size();
        {
            size_t no = (size_t)n_out;
            for (size_t o = 0; o < no; ++o)
                for (size_t i0 = 0; i0 < (size_t)n_in; ++i0) {
                    size_t a = state.expr[i0].rows[o].size();
                    if (!a)
                        continue;
                    tr->r0 += (int64_t)(a * (a - 1) / 2);
                    for (size_t i1 = i0 + 1; i1 < (size_t)n_in; ++i1)
                        tr->r0 += (int64_t)(a * state.expr[i1].rows[o].size());
                }
        }
        std::string m(method);
        int64_t it = 0;
        double counter_seconds = 0.0;
        while (!state.freq_stat.empty() && m != "dummy") {
            if (max_iters >= 0 && it >= max_iters)
                break;
            // the budget applies to the greedy loop; create_state always runs to completion
            if (time_limit_s > 0 &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - tr->create_seconds - counter_seconds > time_limit_s)
                break;
            int64_t fsz = (int64_t)state.freq_stat.size();
            Pair p = select(state, m);
            if (p.id0 == -1 || p.id1 == -1)
                break;
            tr->f_sizes.push_back(fsz);
            tr->pairs.insert(tr->pairs.end(), {p.id0, p.id1, (int64_t)p.shift, (int64_t)p.sub});
            update_expr(state, p, adder_size, carry_size);
            if (want_counters) { // bookkeeping of the checker, not reference work: excluded from `seconds`
                auto c0 = std::chrono::steady_clock::now();
                tr->r_sizes.push_back(count_raw_pairs(state, p));
                counter_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();
            }
            update_stats(state, p);
            ++it;
        }
        tr->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - counter_seconds;
        for (auto &e : state.expr)
            for (auto &r : e.rows)
                tr->d_final += (int64_t)r.size();
        return tr;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
int64_t ref_trace_len(void *h) { return (int64_t) static_cast<RefTrace *>(h)->f_sizes.size(); }
// scalars: f0, r0, d0, d_final, seconds, create_seconds
void ref_trace_scalars(void *h, double *out) {
    auto *t = static_cast<RefTrace *>(h);
    out[0] = (double)t->f0;
    out[1] = (double)t->r0;
    out[2] = (double)t->d0;
    out[3] = (double)t->d_final;
    out[4] = t->seconds;
    out[5] = t->create_seconds;
}
void ref_trace_copy(void *h, int64_t *pairs, int64_t *f_sizes, int64_t *r_sizes) {
    auto *t = static_cast<RefTrace *>(h);
    std::copy(t->pairs.begin(), t->pairs.end(), pairs);
    std::copy(t->f_sizes.begin(), t->f_sizes.end(), f_sizes);
    if (r_sizes && !t->r_sizes.empty())
        std::copy(t->r_sizes.begin(), t->r_sizes.end(), r_sizes);
}
void ref_trace_free(void *h) { delete static_cast<RefTrace *>(h); }

// ---- initial histogram dump (state_opr.cc:115-144): entries as [n,5] int64 (id0,id1,shift,sub,count)
int64_t ref_freq_init(const float *kernel, int64_t n_in, int64_t n_out, const float *qint, int64_t *out, int64_t cap) {
    std::vector<QInterval> q = to_qints(qint, n_in);
    if (q.empty())
        q.assign(n_in, QInterval{-128.0, 127.0, 1.0});
    std::vector<float> l(n_in, 0.0f);
    DAState state = create_state(to_xarr(kernel, n_in, n_out), q, l);
    int64_t n = 0;
    for (auto &kv : state.freq_stat) {
        if (n < cap) {
            out[5 * n + 0] = kv.first.id0;
            out[5 * n + 1] = kv.first.id1;
            out[5 * n + 2] = kv.first.shift;
            out[5 * n + 3] = kv.first.sub;
            out[5 * n + 4] = kv.second;
        }
        ++n;
    }
    return n;
}

} // extern "C"
