// cmvm_lib.cu -- host driver and C ABI (include/da4ml_b200_cmvm.h) of the B200-native CMVM solver.
//
// Host-side control mirrors the reference's api.cc: `solve` (candidate search over decompose_dc,
// api.cc:147-250) and `_solve` (two-stage driver with the latency-retry loop, api.cc:28-145).  The
// reference parallelises the candidates with OpenMP; here all pending solve_single jobs of all
// candidates (and of all problems of a batch) are solved concurrently by one persistent kernel
// launch per round.  All arithmetic of the path runs in the kernels of cmvm_kernels.cuh /
// cmvm_decompose.cuh; nothing here falls back to the CPU.
#include "../../include/da4ml_b200_cmvm.h"
#include "cmvm_decompose.cuh"
#include "cmvm_kernels.cuh"
#include "dais_replay.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace da {

thread_local std::string g_err;
static std::mutex g_mutex;
static cudaStream_t g_stream = nullptr;
static int g_group_override = 0;
static int g_accounting = 0;

struct ApiError : std::runtime_error {
    int code;
    ApiError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CK(expr)                                                                                                  \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess)                                                                                    \
            throw ApiError(DA4ML_E_CUDA, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);    \
    } while (0)

// grow-only device buffer, reused across calls
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool fresh = false; // true right after (re)allocation: contents were zeroed
    void ensure(size_t bytes, bool zero_on_alloc) {
        fresh = false;
        if (bytes <= cap)
            return;
        if (p)
            CK(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        CK(cudaMalloc(&p, want));
        cap = want;
        if (zero_on_alloc)
            CK(cudaMemsetAsync(p, 0, want, g_stream));
        fresh = true;
    }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap)
            return;
        if (p)
            CK(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
        CK(cudaMallocHost(&p, bytes + bytes / 4 + 4096));
        cap = bytes + bytes / 4 + 4096;
    }
};

struct Carver { // bump allocator over a byte range (256 B aligned pieces)
    size_t off = 0;
    size_t take(size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~size_t(255);
        return o;
    }
};

static DevBuf g_job_arena, g_ws_arena, g_slab_arena, g_desc_arena;
static PinBuf g_pin_up, g_pin_down;
static int g_sm_count = 0, g_max_coop = 0;

struct Timing {
    double device_ms = 0;
    double solve_ms = 0;       // time inside cmvm_solve_kernel launches
    int64_t launches = 0;
    int64_t solve_launches = 0;
    double algo_bytes = 0;     // algorithmic bytes of every solve_single executed
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    std::vector<char> is_solve;
    void mark_solve() { is_solve.back() = 1; }
    void begin() {
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        CK(cudaEventRecord(a, g_stream));
        pending.push_back({a, b});
        is_solve.push_back(0);
    }
    void end(int n_launch) {
        CK(cudaEventRecord(pending.back().second, g_stream));
        launches += n_launch;
    }
    void collect() { // call after a stream sync
        for (size_t i = 0; i < pending.size(); ++i) {
            auto &pr = pending[i];
            float ms = 0;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
                device_ms += ms;
                if (is_solve[i])
                    solve_ms += ms;
            }
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        pending.clear();
        is_solve.clear();
    }
};

static void init_device() {
    if (g_sm_count)
        return;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw ApiError(DA4ML_E_CUDA, "no CUDA device available (the CMVM solver has no CPU fallback)");
    int dev = 0;
    CK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    int per_sm = 0;
    CK(cudaFuncSetAttribute(cmvm_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    CK(cudaFuncSetAttribute(cmvm_solve_kernel_x2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cmvm_solve_kernel, 512, 216 * 1024));
    if (per_sm < 1)
        throw ApiError(DA4ML_E_CUDA, "cmvm_solve_kernel cannot be made resident on this device");
    g_sm_count = prop.multiProcessorCount;
    g_max_coop = g_sm_count; // one persistent CTA per SM
}

static int parse_method(const std::string &m) {
    if (m == "mc")
        return M_MC;
    if (m == "mc-dc")
        return M_MC_DC;
    if (m == "mc-pdc")
        return M_MC_PDC;
    if (m == "wmc")
        return M_WMC;
    if (m == "wmc-dc")
        return M_WMC_DC;
    if (m == "wmc-pdc")
        return M_WMC_PDC;
    if (m == "dummy")
        return M_DUMMY;
    throw ApiError(DA4ML_E_RUNTIME, "Unknown method: " + m); // cmvm_core.cc:63
}

// ------------------------------------------------------------------------------------------------
// one CSE stage = one solve_single job

struct StageResult {
    int64_t n_in = 0, n_out = 0;
    int carry_size = -1, adder_size = -1;
    std::vector<int64_t> inp_shifts, out_idxs, out_shifts, out_negs;
    std::vector<float4> out_q; // per output: (qmin, qmax, qstep, latency) of its op; (0, 0, inf, 0) for dead outputs
    float cost_sum = 0.0f;     // cost_init + sum of op costs in op order (float, device)
    int64_t n_ops_dev = 0;
    // op records in device layout: they stay on the device until somebody needs them (only the winning candidate's are
    // ever copied back); expanded to the ABI's int64 / float32 tables when a caller asks for them
    const int4 *d_misc = nullptr;
    const float4 *d_q = nullptr;
    const float *d_cost = nullptr;
    bool have_ops = false;
    std::vector<int4> op_misc;   // id0, id1, opcode, data
    std::vector<float4> op_q;    // qmin, qmax, qstep, latency
    std::vector<float> op_cost;
    int64_t counters[32] = {0};
    int64_t n_ops() const { return n_ops_dev; }
};

struct StageJob {
    int n_in = 0, n_out = 0, method = M_WMC, adder_size = -1, carry_size = -1;
    const float *d_kernel = nullptr; // device, [n_in][n_out]
    std::vector<float> qint, lat;    // host
    float cost_init = 0.0f; // float cost accumulated by the earlier stage(s) of the same candidate
    int32_t *trace = nullptr;
    int64_t trace_cap = 0;
    StageResult res;
    // capacity escalation after an overflow status
    bool full_expr = false, global_lists = false;
    int f_mul = 1, t_mul = 1, list_mul = 2;
};

// Device memory for per-job outputs that must outlive one run_stage_jobs call (op tables of every candidate until the
// winner is known): bump allocation over a list of chunks, recycled by the next API call.
struct OutArena {
    std::vector<DevBuf> chunks;
    size_t cur = 0, off = 0;
    void reset() {
        cur = 0;
        off = 0;
    }
    char *take(size_t bytes) {
        bytes = (bytes + 255) & ~size_t(255);
        while (true) {
            if (cur < chunks.size() && off + bytes <= chunks[cur].cap) {
                char *p = (char *)chunks[cur].p + off;
                off += bytes;
                return p;
            }
            if (cur < chunks.size() && off == 0 && chunks[cur].cap < bytes) { // empty chunk too small: regrow it
                chunks[cur].ensure(bytes, false);
                continue;
            }
            if (cur + 1 < chunks.size() || (cur < chunks.size() && off > 0)) {
                if (cur + 1 >= chunks.size())
                    chunks.emplace_back();
                ++cur;
                off = 0;
                if (chunks[cur].cap < bytes)
                    chunks[cur].ensure(std::max<size_t>(bytes, size_t(256) << 20), false);
                continue;
            }
            chunks.emplace_back();
            cur = chunks.size() - 1;
            off = 0;
            chunks[cur].ensure(std::max<size_t>(bytes, size_t(256) << 20), false);
        }
    }
};
static OutArena g_out_arena2;

// copy one stage's op table back (device layout)
static void fetch_ops(StageResult &r) {
    if (r.have_ops)
        return;
    const size_t n = (size_t)r.n_ops_dev;
    r.op_misc.resize(n);
    r.op_q.resize(n);
    r.op_cost.resize(n);
    if (n) {
        CK(cudaMemcpyAsync(r.op_misc.data(), r.d_misc, sizeof(int4) * n, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaMemcpyAsync(r.op_q.data(), r.d_q, sizeof(float4) * n, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaMemcpyAsync(r.op_cost.data(), r.d_cost, sizeof(float) * n, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
    }
    r.have_ops = true;
}

static int ilog2_ceil(int v) {
    int l = 0;
    while ((1 << l) < v)
        ++l;
    return l;
}

// Solve all jobs concurrently (one prep launch + one persistent solve launch per attempt).
static void run_stage_jobs(std::vector<StageJob *> &jobs, Timing &tm, bool want_ops) {
    if (jobs.empty())
        return;
    init_device();
    const int nj = (int)jobs.size();
    std::vector<StageJob *> todo(jobs.begin(), jobs.end());
    for (int attempt = 0; attempt < 6 && !todo.empty(); ++attempt) {
        const int n = (int)todo.size();
        // ---- job arena: inputs + prep outputs
        Carver cj;
        struct Off {
            size_t qint, lat, masks, s0, s1, cold, pmeta;
        };
        std::vector<Off> off(n);
        for (int i = 0; i < n; ++i) {
            StageJob &j = *todo[i];
            off[i].qint = cj.take(sizeof(float) * 3 * j.n_in);
            off[i].lat = cj.take(sizeof(float) * j.n_in);
            off[i].masks = cj.take(sizeof(uint2) * (size_t)j.n_in * j.n_out);
            off[i].s0 = cj.take(j.n_in);
            off[i].s1 = cj.take(j.n_out);
            off[i].cold = cj.take(sizeof(int) * j.n_out);
            off[i].pmeta = cj.take(sizeof(int) * PM_WORDS);
        }
        const size_t job_in_bytes = cj.off;
        g_job_arena.ensure(job_in_bytes, false);
        char *ja = (char *)g_job_arena.p;
        // upload qint/lat through one pinned staging buffer
        {
            size_t up = 0;
            for (int i = 0; i < n; ++i)
                up += sizeof(float) * 4 * todo[i]->n_in;
            g_pin_up.ensure(up + sizeof(ProblemDesc) * n);
            char *hp = (char *)g_pin_up.p;
            size_t o = 0;
            for (int i = 0; i < n; ++i) {
                StageJob &j = *todo[i];
                memcpy(hp + o, j.qint.data(), sizeof(float) * 3 * j.n_in);
                CK(cudaMemcpyAsync(ja + off[i].qint, hp + o, sizeof(float) * 3 * j.n_in, cudaMemcpyHostToDevice, g_stream));
                o += sizeof(float) * 3 * j.n_in;
                memcpy(hp + o, j.lat.data(), sizeof(float) * j.n_in);
                CK(cudaMemcpyAsync(ja + off[i].lat, hp + o, sizeof(float) * j.n_in, cudaMemcpyHostToDevice, g_stream));
                o += sizeof(float) * j.n_in;
            }
        }
        std::vector<ProblemDesc> desc(n);
        for (int i = 0; i < n; ++i) {
            StageJob &j = *todo[i];
            ProblemDesc &d = desc[i];
            memset(&d, 0, sizeof(d));
            d.n_in = j.n_in;
            d.n_out = j.n_out;
            d.method = j.method;
            d.adder_size = j.adder_size;
            d.carry_size = j.carry_size;
            d.kernel = j.d_kernel;
            d.qint = (const float *)(ja + off[i].qint);
            d.lat = (const float *)(ja + off[i].lat);
            d.masks0 = (uint2 *)(ja + off[i].masks);
            d.shift0 = (int8_t *)(ja + off[i].s0);
            d.shift1 = (int8_t *)(ja + off[i].s1);
            d.col_digits = (int *)(ja + off[i].cold);
            d.prep_meta = (int *)(ja + off[i].pmeta);
        }
        g_desc_arena.ensure(sizeof(ProblemDesc) * n + sizeof(GroupWs) * 256 + 4096, false);
        ProblemDesc *d_desc = (ProblemDesc *)g_desc_arena.p;
        CK(cudaMemcpyAsync(d_desc, desc.data(), sizeof(ProblemDesc) * n, cudaMemcpyHostToDevice, g_stream));
        tm.begin();
        cmvm_prep_kernel<<<n, 256, 0, g_stream>>>(d_desc);
        tm.end(1);
        CK(cudaGetLastError());
        std::vector<int> pmeta((size_t)n * PM_WORDS);
        for (int i = 0; i < n; ++i)
            CK(cudaMemcpyAsync(&pmeta[(size_t)i * PM_WORDS], desc[i].prep_meta, sizeof(int) * PM_WORDS, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        tm.collect();

        // ---- capacities, group geometry
        bool accounting = g_accounting != 0;
        for (int i = 0; i < n; ++i)
            accounting = accounting || todo[i]->trace_cap > 0;
        const char *env_acc = getenv("DA4ML_B200_ACCOUNTING");
        if (env_acc && atoi(env_acc) > 0)
            accounting = true;
        // two 256-thread CTAs per SM (opt-in), one 512-thread CTA per SM otherwise
        bool x2 = false; // measured (round 1): no gain for the 256x256 default solve, 8 % at 128x128; opt-in via DA4ML_B200_CTA_THREADS=256
        if (const char *ev = getenv("DA4ML_B200_CTA_THREADS"))
            x2 = atoi(ev) == 256;
        const int coop = x2 ? 2 * g_max_coop : g_max_coop;
        const int cta_threads = x2 ? 256 : 512;
        // ---- per-job quantities that do not depend on the group size
        Carver co;
        co.off = job_in_bytes;
        struct OOff {
            size_t misc, q, cost, oi, os, on, meta, trace;
        };
        std::vector<OOff> oo(n);
        long long max_cols = 0, max_colcap = 0, max_slab = 0, max_heap = 0, max_ecap = 0, max_rows = 0, want = 1;
        bool force_global_lists = false;
        for (int i = 0; i < n; ++i) {
            StageJob &j = *todo[i];
            const int *pm = &pmeta[(size_t)i * PM_WORDS];
            ProblemDesc &d = desc[i];
            const long long d0 = pm[PM_D0];
            d.nbits = pm[PM_NBITS];
            d.log_s = std::max(1, ilog2_ceil(2 * (2 * d.nbits - 1)));
            long long t_cap = j.full_expr ? d0 : std::min<long long>(d0, d0 / 2 + 1024);
            d.e_cap = (int)(j.n_in + t_cap + 1);
            d.ops_cap = (int)(j.n_in + d0 + 1);
            d.col_cap = pm[PM_COLCAP] + 1;
            d.heap_lane_cap = (std::min(d.nbits, 32) + 1) * ((d.col_cap + 31) / 32) + 2;
            if (((long long)3 * d.e_cap << d.log_s) >= (1LL << 32) || d.e_cap >= (1 << 28))
                throw ApiError(DA4ML_E_CAPACITY, "problem too large for 32-bit counter indices");
            d.op_misc = (int4 *)g_out_arena2.take(sizeof(int4) * d.ops_cap);
            d.op_q = (float4 *)g_out_arena2.take(sizeof(float4) * d.ops_cap);
            d.op_cost = (float *)g_out_arena2.take(sizeof(float) * d.ops_cap);
            d.out_q = (float4 *)g_out_arena2.take(sizeof(float4) * j.n_out);
            d.cost_init = j.cost_init;
            oo[i].oi = co.take(sizeof(int) * j.n_out);
            oo[i].os = co.take(sizeof(int) * j.n_out);
            oo[i].on = co.take(sizeof(int) * j.n_out);
            oo[i].meta = co.take(sizeof(long long) * META_WORDS);
            d.trace_cap = (int)std::min<long long>(j.trace_cap, t_cap + 1);
            oo[i].trace = co.take(sizeof(int) * 5 * (size_t)std::max(d.trace_cap, 1));
            max_cols = std::max<long long>(max_cols, j.n_out);
            max_colcap = std::max<long long>(max_colcap, d.col_cap);
            max_rows = std::max<long long>(max_rows, j.n_in);
            max_ecap = std::max<long long>(max_ecap, d.e_cap);
            max_heap = std::max<long long>(max_heap, (long long)j.n_out * 32 * d.heap_lane_cap);
            max_slab = std::max<long long>(max_slab, (long long)3 * d.e_cap << d.log_s);
            want = std::max(want, std::min<long long>(coop, std::max<long long>(1, d0 / (x2 ? 192 : 384))));
            force_global_lists = force_global_lists || j.global_lists;
        }
        long long list_req = 0; // shortest shared-memory list we accept (the hard bound is col_cap; observed maxima are ~1.6 x n_in)
        for (int i = 0; i < n; ++i)
            list_req = std::max<long long>(list_req, (long long)todo[i]->list_mul * todo[i]->n_in + 64);
        // ---- plan for a given group size: shared-memory layout + per-CTA capacities
        struct Plan {
            LaunchCfg cfg;
            long long max_fcap, max_touch;
            size_t smem_bytes;
        };
        auto plan_for = [&](int G) {
            Plan P;
            memset(&P.cfg, 0, sizeof(P.cfg));
            P.max_fcap = 0;
            P.max_touch = 0;
            for (int i = 0; i < n; ++i) {
                StageJob &j = *todo[i];
                const int *pm = &pmeta[(size_t)i * PM_WORDS];
                const long long d0 = pm[PM_D0];
                long long fcap_total = (128 * d0 + 65536) * j.f_mul;
                P.max_fcap = std::max(P.max_fcap, fcap_total / G + fcap_total / (2 * G) + 8192);
                long long cols_per_cta = (j.n_out + G - 1) / G;
                long long touch = cols_per_cta * 3 * std::min(desc[i].nbits, 32) * (long long)pm[PM_DCOL_MAX] / 4 * j.t_mul + 4096;
                P.max_touch = std::max(P.max_touch, touch);
            }
            if (P.max_fcap >= (1LL << 27))
                P.max_fcap = (1LL << 27) - 1;
            LaunchCfg &cfg = P.cfg;
            cfg.G = G;
            cfg.cpc = (int)((max_cols + G - 1) / G);
            cfg.accounting = accounting ? 1 : 0;
            if (const char *ms = getenv("DA4ML_B200_MAX_STEPS"))
                cfg.max_steps = atoi(ms); // developer knob (results are then incomplete)
            const long long budget = x2 ? 96 * 1024 : 212 * 1024;
            cfg.chunk_log = 6;
            while ((((P.max_fcap >> cfg.chunk_log) + 2) * 17) > (x2 ? 28 : 56) * 1024)
                ++cfg.chunk_log;
            cfg.nchunk_cap = (int)((P.max_fcap >> cfg.chunk_log) + 2);
            cfg.touch_smem = 0; // touched counters live in global memory: every CTA of the group harvests a share
            long long used = (long long)cfg.nchunk_cap * 17 + (long long)cfg.cpc * (4 + (long long)sizeof(ActCol)) + 64;
            long long lcap = (budget - used) / (12LL * cfg.cpc);
            if (lcap >= max_colcap)
                lcap = max_colcap;
            else if (lcap < std::min<long long>(max_colcap, list_req))
                lcap = 0; // too short to be safe: a larger group is tried first, else the lists stay in global memory
            if (force_global_lists || getenv("DA4ML_B200_GLOBAL_LISTS"))
                lcap = 0;
            cfg.lcap = (int)lcap;
            P.smem_bytes = (size_t)cfg.nchunk_cap * 17 + (size_t)cfg.cpc * (4 + sizeof(ActCol)) + 12ull * cfg.cpc * cfg.lcap + 64;
            return P;
        };
        // group size: as many concurrent problems as possible, but never so few CTAs per problem that its column
        // lists fall out of shared memory -- the jobs then run in waves over coop / G groups
        int G = (int)std::min<long long>(want, std::max(1, coop / n));
        if (!force_global_lists && !getenv("DA4ML_B200_GLOBAL_LISTS")) {
            while (G < std::min<long long>(want, coop) && plan_for(G).cfg.lcap == 0)
                ++G;
            // equal waves: with `waves` passes over coop / G groups, spread the CTAs over ceil(n / waves) groups
            const int waves = (n + (coop / G) - 1) / (coop / G);
            const int groups = (n + waves - 1) / waves;
            G = (int)std::min<long long>(want, std::max(G, coop / groups));
        }
        if (g_group_override > 0)
            G = std::min(g_group_override, coop);
        if (const char *env = getenv("DA4ML_B200_GROUP"))
            if (atoi(env) > 0)
                G = std::min(atoi(env), coop);
        const int n_groups = std::max(1, std::min(n, coop / G));
        const Plan plan = plan_for(G);
        const LaunchCfg cfg = plan.cfg;
        const long long max_fcap = plan.max_fcap, max_touch = plan.max_touch;
        const size_t smem_bytes = plan.smem_bytes;
        static DevBuf g_out_arena;
        g_out_arena.ensure(co.off - job_in_bytes, false);
        char *oa = (char *)g_out_arena.p - job_in_bytes;
        for (int i = 0; i < n; ++i) {
            ProblemDesc &d = desc[i];
            d.out_idx = (int *)(oa + oo[i].oi);
            d.out_shift = (int *)(oa + oo[i].os);
            d.out_neg = (int *)(oa + oo[i].on);
            d.result_meta = (long long *)(oa + oo[i].meta);
            d.trace = todo[i]->trace_cap > 0 ? (int *)(oa + oo[i].trace) : nullptr;
            CK(cudaMemsetAsync(d.result_meta, 0, sizeof(long long) * META_WORDS, g_stream));
        }
        // ---- group workspaces
        Carver cw;
        struct WOff {
            size_t ents, len, colk, mod, fseg, touch, slots, heap, bar, xchg;
        };
        std::vector<WOff> wo(n_groups);
        for (int gi = 0; gi < n_groups; ++gi) {
            wo[gi].ents = cw.take(cfg.lcap > 0 ? 256 : sizeof(uint32_t) * 3 * max_cols * max_colcap);
            wo[gi].len = cw.take(sizeof(int) * max_cols);
            wo[gi].colk = cw.take(sizeof(int) * max_cols);
            wo[gi].mod = cw.take(sizeof(uint32_t) * max_ecap);
            wo[gi].fseg = cw.take(sizeof(FEnt) * (size_t)G * max_fcap);
            wo[gi].touch = cw.take(sizeof(uint32_t) * (size_t)G * max_touch);
            wo[gi].slots = cw.take(sizeof(uint4) * 2 * G);
            wo[gi].heap = cw.take(sizeof(uint4) * 2 * max_heap);
            wo[gi].bar = cw.take(256);
            wo[gi].xchg = cw.take(sizeof(unsigned long long) * 2 * 4 * G);
        }
        g_ws_arena.ensure(cw.off, false);
        const size_t slab_bytes_each = ((size_t)max_slab * sizeof(uint32_t) + 255) & ~size_t(255);
        g_slab_arena.ensure(slab_bytes_each * n_groups, true); // counters must start (and are left) zero
        std::vector<GroupWs> gws(n_groups);
        char *wa = (char *)g_ws_arena.p;
        for (int gi = 0; gi < n_groups; ++gi) {
            GroupWs &w = gws[gi];
            w.col_u32 = (uint32_t *)(wa + wo[gi].ents);
            w.col_len = (int *)(wa + wo[gi].len);
            w.col_k = (int *)(wa + wo[gi].colk);
            w.slab = (uint32_t *)((char *)g_slab_arena.p + slab_bytes_each * gi);
            w.mod_step = (uint32_t *)(wa + wo[gi].mod);
            w.fseg = (FEnt *)(wa + wo[gi].fseg);
            w.touch = (uint32_t *)(wa + wo[gi].touch);
            w.slots = (uint4 *)(wa + wo[gi].slots);
            w.heap = (uint4 *)(wa + wo[gi].heap);
            w.barrier = (unsigned *)(wa + wo[gi].bar);
            w.fseg_cap = (int)max_fcap;
            w.touch_cap = (int)max_touch;
            w.heap_cap = max_heap;
            w.xchg = (unsigned long long *)(wa + wo[gi].xchg);
            CK(cudaMemsetAsync(w.barrier, 0, 256, g_stream));
            CK(cudaMemsetAsync(w.xchg, 0, sizeof(unsigned long long) * 2 * 4 * G, g_stream));
        }
        GroupWs *d_gws = (GroupWs *)((char *)g_desc_arena.p + ((sizeof(ProblemDesc) * n + 255) & ~size_t(255)));
        // biggest problems first so that the groups finish together
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i)
            order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pmeta[(size_t)a * PM_WORDS + PM_D0] > pmeta[(size_t)b * PM_WORDS + PM_D0]; });
        std::vector<ProblemDesc> sorted(n);
        for (int i = 0; i < n; ++i)
            sorted[i] = desc[order[i]];
        {
            g_pin_up.ensure(sizeof(ProblemDesc) * n + sizeof(GroupWs) * n_groups);
            char *hp = (char *)g_pin_up.p; // (the earlier uploads from this buffer have completed: the stream was synced)
            memcpy(hp, sorted.data(), sizeof(ProblemDesc) * n);
            memcpy(hp + sizeof(ProblemDesc) * n, gws.data(), sizeof(GroupWs) * n_groups);
            CK(cudaMemcpyAsync(d_desc, hp, sizeof(ProblemDesc) * n, cudaMemcpyHostToDevice, g_stream));
            CK(cudaMemcpyAsync(d_gws, hp + sizeof(ProblemDesc) * n, sizeof(GroupWs) * n_groups, cudaMemcpyHostToDevice, g_stream));
        }
        // ---- persistent solve kernel (cooperative launch: all CTAs must be co-resident for the group barriers)
        {
            const ProblemDesc *a0 = d_desc;
            int a1 = n;
            const GroupWs *a2 = d_gws;
            LaunchCfg a3 = cfg;
            void *args[] = {(void *)&a0, (void *)&a1, (void *)&a2, (void *)&a3};
            tm.begin();
            CK(cudaLaunchCooperativeKernel(x2 ? (void *)cmvm_solve_kernel_x2 : (void *)cmvm_solve_kernel, dim3(n_groups * G), dim3(cta_threads), args, smem_bytes, g_stream));
            tm.end(1);
            tm.solve_launches += 1;
            tm.mark_solve();
        }
        // ---- results
        std::vector<long long> meta((size_t)n * META_WORDS);
        for (int i = 0; i < n; ++i)
            CK(cudaMemcpyAsync(&meta[(size_t)i * META_WORDS], desc[i].result_meta, sizeof(long long) * META_WORDS, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        tm.collect();
        std::vector<StageJob *> again;
        bool dirty_slab = false;
        size_t down = 0;
        for (int i = 0; i < n; ++i) {
            const long long *m = &meta[(size_t)i * META_WORDS];
            if (m[META_STATUS] == ST_OK)
                down += (want_ops ? (size_t)m[META_N_OPS] * 36 : 0) + (size_t)todo[i]->n_out * 28 + todo[i]->n_in + todo[i]->n_out + 5 * 4 * (size_t)desc[i].trace_cap + 1024;
        }
        g_pin_down.ensure(down + 4096);
        char *dp = (char *)g_pin_down.p;
        struct DOff {
            size_t misc, q, cost, oi, os, on, oq, s0, tr;
        };
        std::vector<DOff> dof(n);
        size_t dofs = 0;
        auto dtake = [&](size_t b) {
            size_t o = dofs;
            dofs += (b + 15) & ~size_t(15);
            return o;
        };
        for (int i = 0; i < n; ++i) {
            StageJob &j = *todo[i];
            const long long *m = &meta[(size_t)i * META_WORDS];
            if (m[META_STATUS] != ST_OK) {
                dirty_slab = true;
                switch ((int)m[META_STATUS]) {
                case ST_EXPR_OVERFLOW:
                    if (j.full_expr)
                        throw ApiError(DA4ML_E_CAPACITY, "expression table overflow");
                    j.full_expr = true;
                    break;
                case ST_FSEG_OVERFLOW:
                    j.f_mul *= 4;
                    break;
                case ST_TOUCH_OVERFLOW:
                    j.t_mul *= 4;
                    break;
                case ST_LIST_OVERFLOW:
                    if (j.global_lists)
                        throw ApiError(DA4ML_E_CAPACITY, "column list overflow");
                    if (j.list_mul < 16)
                        j.list_mul *= 2; // ask for longer shared-memory lists (i.e. more CTAs per problem) first
                    else
                        j.global_lists = true;
                    break;
                default:
                    throw ApiError(DA4ML_E_CAPACITY, "internal capacity overflow, status " + std::to_string((int)m[META_STATUS]));
                }
                again.push_back(&j);
                continue;
            }
            const size_t n_ops = (size_t)m[META_N_OPS];
            if (want_ops) {
                dof[i].misc = dtake(sizeof(int4) * n_ops);
                dof[i].q = dtake(sizeof(float4) * n_ops);
                dof[i].cost = dtake(sizeof(float) * n_ops);
            }
            dof[i].oq = dtake(sizeof(float4) * j.n_out);
            dof[i].oi = dtake(sizeof(int) * j.n_out);
            dof[i].os = dtake(sizeof(int) * j.n_out);
            dof[i].on = dtake(sizeof(int) * j.n_out);
            dof[i].s0 = dtake(j.n_in);
            dof[i].tr = dtake(sizeof(int) * 5 * (size_t)std::max(desc[i].trace_cap, 1));
            if (want_ops) {
                CK(cudaMemcpyAsync(dp + dof[i].misc, desc[i].op_misc, sizeof(int4) * n_ops, cudaMemcpyDeviceToHost, g_stream));
                CK(cudaMemcpyAsync(dp + dof[i].q, desc[i].op_q, sizeof(float4) * n_ops, cudaMemcpyDeviceToHost, g_stream));
                CK(cudaMemcpyAsync(dp + dof[i].cost, desc[i].op_cost, sizeof(float) * n_ops, cudaMemcpyDeviceToHost, g_stream));
            }
            CK(cudaMemcpyAsync(dp + dof[i].oq, desc[i].out_q, sizeof(float4) * j.n_out, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaMemcpyAsync(dp + dof[i].oi, desc[i].out_idx, sizeof(int) * j.n_out, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaMemcpyAsync(dp + dof[i].os, desc[i].out_shift, sizeof(int) * j.n_out, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaMemcpyAsync(dp + dof[i].on, desc[i].out_neg, sizeof(int) * j.n_out, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaMemcpyAsync(dp + dof[i].s0, desc[i].shift0, j.n_in, cudaMemcpyDeviceToHost, g_stream));
            if (desc[i].trace)
                CK(cudaMemcpyAsync(dp + dof[i].tr, desc[i].trace, sizeof(int) * 5 * (size_t)desc[i].trace_cap, cudaMemcpyDeviceToHost, g_stream));
        }
        CK(cudaStreamSynchronize(g_stream));
        for (int i = 0; i < n; ++i) {
            StageJob &j = *todo[i];
            const long long *m = &meta[(size_t)i * META_WORDS];
            if (m[META_STATUS] != ST_OK)
                continue;
            StageResult &r = j.res;
            const size_t n_ops = (size_t)m[META_N_OPS];
            r.n_in = j.n_in;
            r.n_out = j.n_out;
            r.carry_size = j.carry_size;
            r.adder_size = j.adder_size;
            for (int w = 0; w < META_WORDS; ++w)
                r.counters[w] = m[w];
            // algorithmic bytes of this solve_single; sum_F is exact only in accounting mode
            tm.algo_bytes += 8.0 * j.n_in * j.n_out + 12.0 * (double)m[META_F0] + 2.0 * (double)m[META_R0] + 12.0 * (double)m[META_SUM_F] +
                             10.0 * (double)m[META_SUM_R] + 2.0 * (double)m[META_D_FINAL] + 56.0 * (double)m[META_N_OPS];
            r.counters[10] = pmeta[(size_t)i * PM_WORDS + PM_D0];
            r.counters[11] = pmeta[(size_t)i * PM_WORDS + PM_NBITS];
            r.counters[12] = G;
            r.counters[15] = cfg.lcap;
            r.counters[12] = G;
            r.inp_shifts.resize(j.n_in);
            const int8_t *s0 = (const int8_t *)(dp + dof[i].s0);
            for (int k = 0; k < j.n_in; ++k)
                r.inp_shifts[k] = s0[k];
            r.out_idxs.resize(j.n_out);
            r.out_shifts.resize(j.n_out);
            r.out_negs.resize(j.n_out);
            const int *oi = (const int *)(dp + dof[i].oi), *os = (const int *)(dp + dof[i].os), *on = (const int *)(dp + dof[i].on);
            for (int k = 0; k < j.n_out; ++k) {
                r.out_idxs[k] = oi[k];
                r.out_shifts[k] = os[k];
                r.out_negs[k] = on[k];
            }
            r.n_ops_dev = (int64_t)n_ops;
            r.d_misc = desc[i].op_misc;
            r.d_q = desc[i].op_q;
            r.d_cost = desc[i].op_cost;
            r.have_ops = false;
            r.out_q.assign((const float4 *)(dp + dof[i].oq), (const float4 *)(dp + dof[i].oq) + j.n_out);
            {
                const uint32_t bits = (uint32_t)m[META_COST_BITS];
                memcpy(&r.cost_sum, &bits, 4);
            }
            if (want_ops) {
                r.op_misc.assign((const int4 *)(dp + dof[i].misc), (const int4 *)(dp + dof[i].misc) + n_ops);
                r.op_q.assign((const float4 *)(dp + dof[i].q), (const float4 *)(dp + dof[i].q) + n_ops);
                r.op_cost.assign((const float *)(dp + dof[i].cost), (const float *)(dp + dof[i].cost) + n_ops);
                r.have_ops = true;
            }
            if (j.trace && j.trace_cap > 0) {
                int64_t rows = std::min<int64_t>(std::min<int64_t>(j.trace_cap, desc[i].trace_cap), m[META_T]);
                memcpy(j.trace, dp + dof[i].tr, sizeof(int) * 5 * (size_t)rows);
            }
        }
        if (dirty_slab)
            CK(cudaMemsetAsync(g_slab_arena.p, 0, g_slab_arena.cap, g_stream));
        todo.swap(again);
    }
    if (!todo.empty())
        throw ApiError(DA4ML_E_CAPACITY, "could not size the solver buffers after repeated attempts");
    (void)nj;
}

// ------------------------------------------------------------------------------------------------
// kernel_decompose on the device: results stay on the device for the stage jobs

// ------------------------------------------------------------------------------------------------
// _solve state machine (api.cc:28-145), one per decompose_dc candidate

struct PipelineImpl {
    std::vector<StageResult> stages;
    double device_ms = 0, solve_ms = 0, algo_bytes = 0;
    int64_t launches = 0, solve_launches = 0;
};

struct Candidate {
    // fixed
    int problem = 0;
    std::string method0, method1;
    int hard_dc = -1;
    int decompose_dc = -2; // current value (after the api.cc:74-80 clamp)
    int adder_size = -1, carry_size = -1;
    float latency_allowed = std::numeric_limits<float>::infinity();
    // state
    int phase = 0; // 0: needs stage 0, 1: needs stage 1, 2: done
    StageJob job0, job1;
    float *d_m0 = nullptr, *d_m1 = nullptr;
    int *d_map = nullptr;
};

struct Problem {
    const float *h_kernel = nullptr;
    int n_in = 0, n_out = 0;
    std::vector<float> qint, lat;
    // device
    float *d_kernel = nullptr, *d_aug = nullptr;
    int *d_dist = nullptr;
    int8_t *d_sign = nullptr, *d_s0 = nullptr, *d_s1 = nullptr;
    bool need_min_lat = false;
    StageJob min_lat_job;
    float min_lat = std::numeric_limits<float>::infinity();
    std::vector<int> cand; // indices into the candidate vector
};

static bool ends_with(const std::string &s, const std::string &suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

static float stage_max_latency(const StageResult &r) {
    float m = 0.0f;
    for (size_t k = 0; k < r.out_idxs.size(); ++k) {
        float lat = r.out_idxs[k] >= 0 ? r.out_q[k].w : 0.0f;
        m = std::max(m, lat);
    }
    return m;
}

static void solve_many(
    int64_t n_problems, const float *const *kernels, const int64_t *n_in, const int64_t *n_out, const std::string &method0_in,
    const std::string &method1_in, int hard_dc, int decompose_dc, const float *const *qints, const float *const *lats,
    int adder_size, int carry_size, bool search_all, std::vector<std::unique_ptr<PipelineImpl>> &out, bool kernels_on_device = false
) {
    init_device();
    g_out_arena2.reset();
    Timing tm;
    // validate methods up front (the reference throws from the worker, api.cc:231-240)
    parse_method(method0_in);
    if (method1_in != "auto")
        parse_method(method1_in);

    std::vector<Problem> probs(n_problems);
    std::vector<Candidate> cands;
    // ---- device residency of the inputs + decomposition scratch
    static DevBuf g_base, g_pool;
    Carver cb;
    struct BOff {
        size_t k, aug, dist, sign, s0, s1;
    };
    std::vector<BOff> bo(n_problems);
    size_t up_bytes = 0;
    for (int64_t pi = 0; pi < n_problems; ++pi) {
        Problem &P = probs[pi];
        if (n_in[pi] <= 0 || n_out[pi] <= 0)
            throw ApiError(DA4ML_E_INVALID, "kernel must be a non-empty 2D array");
        P.n_in = (int)n_in[pi];
        P.n_out = (int)n_out[pi];
        P.h_kernel = kernels[pi];
        const float *q = qints ? qints[pi] : nullptr;
        const float *l = lats ? lats[pi] : nullptr;
        P.qint.resize(3 * (size_t)P.n_in);
        P.lat.resize(P.n_in);
        for (int i = 0; i < P.n_in; ++i) { // api.cc:161-174 defaults
            P.qint[3 * i + 0] = q ? q[3 * i + 0] : -128.0f;
            P.qint[3 * i + 1] = q ? q[3 * i + 1] : 127.0f;
            P.qint[3 * i + 2] = q ? q[3 * i + 2] : 1.0f;
            P.lat[i] = l ? l[i] : 0.0f;
        }
        const size_t n = (size_t)P.n_out + 1;
        bo[pi].k = cb.take(sizeof(float) * (size_t)P.n_in * P.n_out);
        bo[pi].aug = cb.take(sizeof(float) * (size_t)P.n_in * n);
        bo[pi].dist = cb.take(sizeof(int) * n * n);
        bo[pi].sign = cb.take(n * n);
        bo[pi].s0 = cb.take(P.n_in);
        bo[pi].s1 = cb.take(P.n_out);
        up_bytes += sizeof(float) * (size_t)P.n_in * P.n_out;
    }
    g_base.ensure(cb.off, false);
    static PinBuf g_pin_k;
    g_pin_k.ensure(up_bytes);
    {
        size_t o = 0;
        for (int64_t pi = 0; pi < n_problems; ++pi) {
            Problem &P = probs[pi];
            char *b = (char *)g_base.p;
            P.d_kernel = (float *)(b + bo[pi].k);
            P.d_aug = (float *)(b + bo[pi].aug);
            P.d_dist = (int *)(b + bo[pi].dist);
            P.d_sign = (int8_t *)(b + bo[pi].sign);
            P.d_s0 = (int8_t *)(b + bo[pi].s0);
            P.d_s1 = (int8_t *)(b + bo[pi].s1);
            size_t bytes = sizeof(float) * (size_t)P.n_in * P.n_out;
            if (kernels_on_device) // inputs already resident in HBM
                CK(cudaMemcpyAsync(P.d_kernel, P.h_kernel, bytes, cudaMemcpyDeviceToDevice, g_stream));
            else {
                memcpy((char *)g_pin_k.p + o, P.h_kernel, bytes);
                CK(cudaMemcpyAsync(P.d_kernel, (char *)g_pin_k.p + o, bytes, cudaMemcpyHostToDevice, g_stream));
            }
            o += bytes;
        }
    }
    // ---- candidates (api.cc:176-201)
    for (int64_t pi = 0; pi < n_problems; ++pi) {
        Problem &P = probs[pi];
        std::vector<std::pair<int, int>> tries; // (hard_dc passed to _solve, decompose_dc passed to _solve)
        if (!search_all)
            tries.push_back({hard_dc, decompose_dc});
        else {
            int _hard_dc = hard_dc < 0 ? 1000000000 : hard_dc;
            int max_dc = std::min(_hard_dc, (int)std::ceil(std::log2((float)P.n_in)));
            for (int d = -1; d <= max_dc; ++d)
                tries.push_back({_hard_dc, d});
        }
        for (auto &tr : tries) {
            Candidate c;
            c.problem = (int)pi;
            c.method0 = method0_in;
            c.method1 = method1_in;
            c.hard_dc = tr.first;
            c.adder_size = adder_size;
            c.carry_size = carry_size;
            // api.cc:41-51
            if (c.method1 == "auto")
                c.method1 = (c.hard_dc >= 6 || ends_with(c.method0, "dc")) ? c.method0 : c.method0 + "-dc";
            if (c.hard_dc == 0 && !ends_with(c.method0, "dc"))
                c.method0 = c.method0 + "-dc";
            // api.cc:74-80
            int log2_n = (int)std::ceil(std::log2((float)P.n_in));
            c.decompose_dc = tr.second == -2 ? std::min(c.hard_dc, log2_n) : std::min({c.hard_dc, tr.second, log2_n});
            if (c.hard_dc >= 0)
                P.need_min_lat = true;
            P.cand.push_back((int)cands.size());
            cands.push_back(std::move(c));
        }
    }
    // ---- decomposition scratch: one (m0, m1, mapping) triple per candidate
    Carver cp;
    struct POff {
        size_t m0, m1, map;
    };
    std::vector<POff> po(cands.size());
    for (size_t ci = 0; ci < cands.size(); ++ci) {
        Problem &P = probs[cands[ci].problem];
        po[ci].m0 = cp.take(sizeof(float) * (size_t)P.n_in * P.n_out);
        po[ci].m1 = cp.take(sizeof(float) * (size_t)P.n_out * P.n_out);
        po[ci].map = cp.take(sizeof(int) * 2 * (size_t)(P.n_out + 1));
    }
    g_pool.ensure(cp.off + sizeof(DecompJob) * cands.size() + 4096, false);
    for (size_t ci = 0; ci < cands.size(); ++ci) {
        char *b = (char *)g_pool.p;
        cands[ci].d_m0 = (float *)(b + po[ci].m0);
        cands[ci].d_m1 = (float *)(b + po[ci].m1);
        cands[ci].d_map = (int *)(b + po[ci].map);
    }
    DecompJob *d_djobs = (DecompJob *)((char *)g_pool.p + ((cp.off + 255) & ~size_t(255)));
    // centre + all-pairs distance once per problem (mat_decompose.cc:64-93)
    tm.begin();
    int n_l = 0;
    for (auto &P : probs) {
        const int n = P.n_out + 1;
        center_kernel<<<1, 256, 0, g_stream>>>(P.d_kernel, P.n_in, P.n_out, P.d_aug, P.d_s0, P.d_s1);
        long long tot = (long long)n * n;
        dist_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, g_stream>>>(P.d_aug, P.n_in, n, P.d_dist, P.d_sign);
        n_l += 2;
    }
    tm.end(n_l);
    CK(cudaGetLastError());

    // minimal_latency (api.cc:11-26, :68-72): to_solution of the un-optimised state
    std::vector<StageJob *> jobs;
    for (auto &P : probs) {
        if (!P.need_min_lat)
            continue;
        StageJob &j = P.min_lat_job;
        j.n_in = P.n_in;
        j.n_out = P.n_out;
        j.method = M_DUMMY;
        j.adder_size = adder_size;
        j.carry_size = carry_size;
        j.d_kernel = P.d_kernel;
        j.qint = P.qint;
        j.lat = P.lat;
        jobs.push_back(&j);
    }
    run_stage_jobs(jobs, tm, false);
    for (auto &P : probs)
        if (P.need_min_lat)
            P.min_lat = stage_max_latency(P.min_lat_job.res);
    for (auto &c : cands)
        if (c.hard_dc >= 0)
            c.latency_allowed = (float)c.hard_dc + probs[c.problem].min_lat; // api.cc:72

    // ---- rounds: every unfinished candidate contributes its next solve_single
    for (int round = 0; round < 4096; ++round) {
        // (re)decompose the candidates that need a stage-0 solve
        std::vector<int> dec;
        for (size_t ci = 0; ci < cands.size(); ++ci)
            if (cands[ci].phase == 0)
                dec.push_back((int)ci);
        bool any_left = !dec.empty();
        for (auto &c : cands)
            any_left = any_left || c.phase == 1;
        if (!any_left)
            break;
        if (!dec.empty()) {
            // api.cc:84-93: once decompose_dc < 0 under a finite hard_dc, both methods are forced
            for (int ci : dec) {
                Candidate &c = cands[ci];
                if (c.decompose_dc < 0 && c.hard_dc >= 0) {
                    if (c.method0 != "dummy") {
                        c.method0 = "wmc-dc";
                        c.method1 = "wmc-dc";
                    }
                    else {
                        c.method0 = "dummy";
                        c.method1 = "dummy";
                    }
                }
            }
            // group by problem: one launch per problem, one CTA per candidate
            std::vector<DecompJob> dj(cands.size());
            for (int ci : dec)
                dj[ci] = DecompJob{cands[ci].decompose_dc, cands[ci].d_m0, cands[ci].d_m1, cands[ci].d_map};
            static PinBuf pin_dj;
            pin_dj.ensure(sizeof(DecompJob) * cands.size());
            CK(cudaStreamSynchronize(g_stream));
            tm.begin();
            int nl = 0;
            size_t run0 = 0;
            // candidates of one problem are contiguous in `cands`
            std::vector<DecompJob> packed;
            for (auto &P : probs) {
                packed.clear();
                for (int ci : P.cand)
                    if (cands[ci].phase == 0)
                        packed.push_back(dj[ci]);
                if (packed.empty())
                    continue;
                memcpy((char *)pin_dj.p + sizeof(DecompJob) * run0, packed.data(), sizeof(DecompJob) * packed.size());
                CK(cudaMemcpyAsync(d_djobs + run0, (char *)pin_dj.p + sizeof(DecompJob) * run0, sizeof(DecompJob) * packed.size(), cudaMemcpyHostToDevice, g_stream));
                const int n = P.n_out + 1;
                const int threads = std::min(1024, std::max(64, (n + 31) / 32 * 32));
                const size_t smem = (size_t)n * 17 + 64;
                if (smem > 48 * 1024)
                    CK(cudaFuncSetAttribute(mst_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                mst_build_kernel<<<(unsigned)packed.size(), threads, smem, g_stream>>>(P.d_aug, P.d_dist, P.d_sign, P.d_s0, P.d_s1, P.n_in, P.n_out, d_djobs + run0);
                run0 += packed.size();
                ++nl;
            }
            tm.end(nl);
            CK(cudaGetLastError());
        }
        for (auto &c : cands) {
            Problem &P = probs[c.problem];
            if (c.phase == 0) {
                StageJob &j = c.job0;
                j = StageJob();
                j.n_in = P.n_in;
                j.n_out = P.n_out;
                j.method = parse_method(c.method0);
                j.adder_size = c.adder_size;
                j.carry_size = c.carry_size;
                j.d_kernel = c.d_m0;
                j.qint = P.qint;
                j.lat = P.lat;
            }
            else if (c.phase == 1) {
                StageJob &j = c.job1;
                j = StageJob();
                j.n_in = P.n_out;
                j.n_out = P.n_out;
                j.method = parse_method(c.method1);
                j.adder_size = c.adder_size;
                j.carry_size = c.carry_size;
                j.d_kernel = c.d_m1;
                j.cost_init = c.job0.res.cost_sum; // the float cost keeps accumulating across the two stages (api.cc:222-227)
                // api.cc:100-115: stage-1 inputs are the RAW op qint/latency of the stage-0 outputs
                const StageResult &r0 = c.job0.res;
                j.qint.resize(3 * (size_t)P.n_out);
                j.lat.resize(P.n_out);
                for (int k = 0; k < P.n_out; ++k) {
                    int64_t idx = r0.out_idxs[k];
                    if (idx >= 0) {
                        j.qint[3 * k + 0] = r0.out_q[k].x;
                        j.qint[3 * k + 1] = r0.out_q[k].y;
                        j.qint[3 * k + 2] = r0.out_q[k].z;
                        j.lat[k] = r0.out_q[k].w;
                    }
                    else {
                        j.qint[3 * k + 0] = 0.0f;
                        j.qint[3 * k + 1] = 0.0f;
                        j.qint[3 * k + 2] = std::numeric_limits<float>::infinity();
                        j.lat[k] = 0.0f;
                    }
                }
            }
        }
        // stage-0 and stage-1 jobs differ wildly in size: run them as separate launches
        std::vector<StageJob *> big, small;
        for (auto &c : cands) {
            if (c.phase == 0)
                big.push_back(&c.job0);
            else if (c.phase == 1)
                small.push_back(&c.job1);
        }
        run_stage_jobs(big, tm, false);
        run_stage_jobs(small, tm, false);
        for (auto &c : cands) {
            const bool both_wmc_dc = c.method0 == "wmc-dc" && c.method1 == "wmc-dc";
            if (c.phase == 0) {
                float max_lat0 = stage_max_latency(c.job0.res);
                if (max_lat0 > c.latency_allowed && (!both_wmc_dc || c.decompose_dc >= 0)) {
                    c.decompose_dc--; // api.cc:117-122
                    if (c.decompose_dc < -64)
                        throw ApiError(DA4ML_E_RUNTIME, "latency constraint cannot be met");
                    continue;
                }
                c.phase = 1;
            }
            else if (c.phase == 1) {
                float max_lat1 = stage_max_latency(c.job1.res);
                if (max_lat1 > c.latency_allowed && (!both_wmc_dc || c.decompose_dc >= 0)) {
                    c.decompose_dc--; // api.cc:133-138
                    if (c.decompose_dc < -64)
                        throw ApiError(DA4ML_E_RUNTIME, "latency constraint cannot be met");
                    c.phase = 0;
                    continue;
                }
                c.phase = 2;
            }
        }
    }
    // ---- argmin over candidates, first minimum wins (api.cc:243-249); cost summed in float in op order (api.cc:222-227)
    out.clear();
    for (auto &P : probs) {
        int best = -1;
        float best_cost = 0;
        for (int ci : P.cand) {
            Candidate &c = cands[ci];
            const float cost = c.job1.res.cost_sum; // stage-0 sum carried into stage 1 on the device
            if (best < 0 || cost < best_cost) {
                best = ci;
                best_cost = cost;
            }
        }
        fetch_ops(cands[best].job0.res); // only the winner's op tables leave the device
        fetch_ops(cands[best].job1.res);
        auto pl = std::make_unique<PipelineImpl>();
        pl->stages.push_back(std::move(cands[best].job0.res));
        pl->stages.push_back(std::move(cands[best].job1.res));
        out.push_back(std::move(pl));
    }
    for (auto &pl : out) {
        pl->device_ms = tm.device_ms;
        pl->launches = tm.launches;
        pl->solve_ms = tm.solve_ms;
        pl->solve_launches = tm.solve_launches;
        pl->algo_bytes = tm.algo_bytes;
    }
}

// small export kernels for the helper entry points ------------------------------------------------
__global__ void csd_export_kernel(const uint2 *masks, int n, int nbits, int8_t *csd) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n)
        return;
    uint2 m = masks[idx];
    for (int k = 0; k < nbits; ++k)
        csd[(size_t)idx * nbits + k] = (int8_t)(((m.x >> k) & 1) - ((m.y >> k) & 1));
}
// plain digits without centring: the matrix is interpreted as already integer (bit_decompose.cc:22-42)
__global__ void int_csd_kernel(const int *x, int n, int nbits, int8_t *csd) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n)
        return;
    int v = x[idx];
    for (int k = nbits - 1; k >= 0; --k) {
        int p2 = (int)(1u << k);
        int thres = (int)(((long long)p2 * 2) / 3);
        int d = (v > thres) - (v < -thres);
        csd[(size_t)idx * nbits + k] = (int8_t)d;
        v -= p2 * d;
    }
}
__global__ void int_absmax_kernel(const int *x, int n, int *out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int v = idx < n ? abs(x[idx]) : 0;
    for (int off = 16; off > 0; off >>= 1)
        v = max(v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0 && v)
        atomicMax(out, v);
}
__global__ void float_to_int_kernel(const float *x, int n, int *out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n)
        out[idx] = (int)x[idx];
}

} // namespace da

// ================================================================================================
using namespace da;

struct da4ml_pipeline {
    std::unique_ptr<PipelineImpl> impl;
};

template <class F> static int guarded(F &&f) {
    try {
        std::lock_guard<std::mutex> lock(g_mutex);
        f();
        return DA4ML_OK;
    }
    catch (const ApiError &e) {
        g_err = e.what();
        return e.code;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return DA4ML_E_RUNTIME;
    }
}

extern "C" {

const char *da4ml_cmvm_last_error(void) { return g_err.c_str(); }

int da4ml_cmvm_device_info(int32_t out[5]) {
    out[0] = 1;
    out[1] = out[2] = out[3] = out[4] = 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
        cudaGetLastError();
        return DA4ML_OK;
    }
    out[1] = ndev;
    if (ndev > 0) {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
            out[2] = prop.multiProcessorCount;
            out[3] = prop.major;
            out[4] = prop.minor;
        }
    }
    return DA4ML_OK;
}

int da4ml_cmvm_set_stream(void *s) {
    g_stream = (cudaStream_t)s;
    return DA4ML_OK;
}
int da4ml_cmvm_set_group_size(int g) {
    g_group_override = g;
    return DA4ML_OK;
}
// Free every device / pinned buffer the library keeps between calls (they are re-grown on demand).
int da4ml_cmvm_release(void) {
    return guarded([&] {
        for (DevBuf *b : {&g_job_arena, &g_ws_arena, &g_slab_arena, &g_desc_arena}) {
            if (b->p)
                cudaFree(b->p);
            b->p = nullptr;
            b->cap = 0;
        }
        for (DevBuf &b : g_out_arena2.chunks) {
            if (b.p)
                cudaFree(b.p);
            b.p = nullptr;
            b.cap = 0;
        }
        g_out_arena2.chunks.clear();
        g_out_arena2.reset();
        for (PinBuf *b : {&g_pin_up, &g_pin_down}) {
            if (b->p)
                cudaFreeHost(b->p);
            b->p = nullptr;
            b->cap = 0;
        }
    });
}
int da4ml_cmvm_set_accounting(int on) {
    g_accounting = on;
    return DA4ML_OK;
}

int da4ml_cmvm_solve_batch(
    int64_t n_problems, const float *const *kernels, const int64_t *n_in, const int64_t *n_out, const char *method0,
    const char *method1, int hard_dc, int decompose_dc, const float *const *qintervals, const float *const *latencies,
    int adder_size, int carry_size, int search_all, da4ml_pipeline_t **out
) {
    return guarded([&] {
        if (n_problems <= 0 || !kernels || !n_in || !n_out || !out || !method0 || !method1)
            throw ApiError(DA4ML_E_INVALID, "invalid argument");
        std::vector<std::unique_ptr<PipelineImpl>> res;
        solve_many(n_problems, kernels, n_in, n_out, method0, method1, hard_dc, decompose_dc, qintervals, latencies, adder_size, carry_size, search_all != 0, res);
        for (int64_t i = 0; i < n_problems; ++i) {
            out[i] = new da4ml_pipeline();
            out[i]->impl = std::move(res[i]);
        }
    });
}

int da4ml_cmvm_solve_batch_device(
    int64_t n_problems, const float *const *kernels_dev, const int64_t *n_in, const int64_t *n_out, const char *method0,
    const char *method1, int hard_dc, int decompose_dc, const float *const *qintervals, const float *const *latencies,
    int adder_size, int carry_size, int search_all, da4ml_pipeline_t **out
) {
    return guarded([&] {
        if (n_problems <= 0 || !kernels_dev || !n_in || !n_out || !out || !method0 || !method1)
            throw ApiError(DA4ML_E_INVALID, "invalid argument");
        std::vector<std::unique_ptr<PipelineImpl>> res;
        solve_many(n_problems, kernels_dev, n_in, n_out, method0, method1, hard_dc, decompose_dc, qintervals, latencies, adder_size, carry_size, search_all != 0, res, true);
        for (int64_t i = 0; i < n_problems; ++i) {
            out[i] = new da4ml_pipeline();
            out[i]->impl = std::move(res[i]);
        }
    });
}

int da4ml_cmvm_solve(
    const float *kernel, int64_t n_in, int64_t n_out, const char *method0, const char *method1, int hard_dc, int decompose_dc,
    const float *qintervals, const float *latencies, int adder_size, int carry_size, int search_all, da4ml_pipeline_t **out
) {
    const float *ks[1] = {kernel};
    const float *qs[1] = {qintervals};
    const float *ls[1] = {latencies};
    return da4ml_cmvm_solve_batch(1, ks, &n_in, &n_out, method0, method1, hard_dc, decompose_dc, qs, ls, adder_size, carry_size, search_all, out);
}

int da4ml_cmvm_solve_single(
    const float *kernel, int64_t n_in, int64_t n_out, const char *method, const float *qintervals, const float *latencies,
    int adder_size, int carry_size, int32_t *trace, int64_t trace_cap, da4ml_pipeline_t **out
) {
    return guarded([&] {
        if (!kernel || n_in <= 0 || n_out <= 0 || !method || !out)
            throw ApiError(DA4ML_E_INVALID, "invalid argument");
        init_device();
        static DevBuf kbuf;
        static PinBuf kpin;
        size_t bytes = sizeof(float) * (size_t)n_in * n_out;
        kbuf.ensure(bytes, false);
        kpin.ensure(bytes);
        memcpy(kpin.p, kernel, bytes);
        CK(cudaMemcpyAsync(kbuf.p, kpin.p, bytes, cudaMemcpyHostToDevice, g_stream));
        StageJob j;
        j.n_in = (int)n_in;
        j.n_out = (int)n_out;
        j.method = parse_method(method);
        j.adder_size = adder_size;
        j.carry_size = carry_size;
        j.d_kernel = (const float *)kbuf.p;
        j.qint.resize(3 * (size_t)n_in);
        j.lat.resize(n_in);
        for (int64_t i = 0; i < n_in; ++i) { // cmvm_core.cc:18-32 defaults
            j.qint[3 * i + 0] = qintervals ? qintervals[3 * i + 0] : -128.0f;
            j.qint[3 * i + 1] = qintervals ? qintervals[3 * i + 1] : 127.0f;
            j.qint[3 * i + 2] = qintervals ? qintervals[3 * i + 2] : 1.0f;
            j.lat[i] = latencies ? latencies[i] : 0.0f;
        }
        j.trace = trace;
        j.trace_cap = trace ? trace_cap : 0;
        Timing tm;
        std::vector<StageJob *> jobs{&j};
        g_out_arena2.reset();
        run_stage_jobs(jobs, tm, true);
        auto pl = std::make_unique<PipelineImpl>();
        pl->stages.push_back(std::move(j.res));
        pl->device_ms = tm.device_ms;
        pl->launches = tm.launches;
        pl->solve_ms = tm.solve_ms;
        pl->solve_launches = tm.solve_launches;
        pl->algo_bytes = tm.algo_bytes;
        *out = new da4ml_pipeline();
        (*out)->impl = std::move(pl);
    });
}

void da4ml_pipeline_free(da4ml_pipeline_t *p) { delete p; }
int64_t da4ml_pipeline_n_stages(const da4ml_pipeline_t *p) { return p ? (int64_t)p->impl->stages.size() : 0; }
int da4ml_pipeline_stage_meta(const da4ml_pipeline_t *p, int64_t s, int64_t meta[5]) {
    if (!p || s < 0 || s >= (int64_t)p->impl->stages.size())
        return DA4ML_E_INVALID;
    const StageResult &r = p->impl->stages[s];
    meta[0] = r.n_in;
    meta[1] = r.n_out;
    meta[2] = r.n_ops();
    meta[3] = r.carry_size;
    meta[4] = r.adder_size;
    return DA4ML_OK;
}
int da4ml_pipeline_stage_copy(
    const da4ml_pipeline_t *p, int64_t s, int64_t *inp_shifts, int64_t *out_idxs, int64_t *out_shifts, int64_t *out_negs,
    int64_t *ops_i, float *ops_f
) {
    if (!p || s < 0 || s >= (int64_t)p->impl->stages.size())
        return DA4ML_E_INVALID;
    const StageResult &r = p->impl->stages[s];
    if (inp_shifts)
        std::copy(r.inp_shifts.begin(), r.inp_shifts.end(), inp_shifts);
    if (out_idxs)
        std::copy(r.out_idxs.begin(), r.out_idxs.end(), out_idxs);
    if (out_shifts)
        std::copy(r.out_shifts.begin(), r.out_shifts.end(), out_shifts);
    if (out_negs)
        std::copy(r.out_negs.begin(), r.out_negs.end(), out_negs);
    const size_t n_ops = r.op_misc.size();
    if (ops_i)
        for (size_t k = 0; k < n_ops; ++k) {
            ops_i[4 * k + 0] = r.op_misc[k].x;
            ops_i[4 * k + 1] = r.op_misc[k].y;
            ops_i[4 * k + 2] = r.op_misc[k].z;
            ops_i[4 * k + 3] = r.op_misc[k].w;
        }
    if (ops_f)
        for (size_t k = 0; k < n_ops; ++k) {
            ops_f[5 * k + 0] = r.op_q[k].x;
            ops_f[5 * k + 1] = r.op_q[k].y;
            ops_f[5 * k + 2] = r.op_q[k].z;
            ops_f[5 * k + 3] = r.op_q[k].w;
            ops_f[5 * k + 4] = r.op_cost[k];
        }
    return DA4ML_OK;
}
int da4ml_pipeline_stage_counters(const da4ml_pipeline_t *p, int64_t s, int64_t counters[32]) {
    if (!p || s < 0 || s >= (int64_t)p->impl->stages.size())
        return DA4ML_E_INVALID;
    std::copy(p->impl->stages[s].counters, p->impl->stages[s].counters + 32, counters);
    return DA4ML_OK;
}
double da4ml_pipeline_device_ms(const da4ml_pipeline_t *p) { return p ? p->impl->device_ms : 0.0; }
int64_t da4ml_pipeline_launches(const da4ml_pipeline_t *p) { return p ? p->impl->launches : 0; }
int da4ml_pipeline_profile(const da4ml_pipeline_t *p, double out[8]) {
    if (!p)
        return DA4ML_E_INVALID;
    out[0] = p->impl->device_ms;
    out[1] = (double)p->impl->launches;
    out[2] = p->impl->solve_ms;
    out[3] = (double)p->impl->solve_launches;
    out[4] = p->impl->algo_bytes;
    out[5] = out[6] = out[7] = 0.0;
    return DA4ML_OK;
}

// ---- helpers -------------------------------------------------------------------------------------
int da4ml_cmvm_csd_decompose(const float *kernel, int64_t n_in, int64_t n_out, int center, int8_t *csd, int8_t *shift0, int8_t *shift1, int64_t *n_bits) {
    return guarded([&] {
        if (!kernel || n_in <= 0 || n_out <= 0 || !csd || !shift0 || !shift1 || !n_bits)
            throw ApiError(DA4ML_E_INVALID, "csd_decompose only supports 2D arrays."); // bit_decompose.cc:48
        init_device();
        const size_t ne = (size_t)n_in * n_out;
        static DevBuf buf;
        Carver c;
        size_t ok = c.take(sizeof(float) * ne), oq = c.take(sizeof(float) * 3 * n_in), ol = c.take(sizeof(float) * n_in), om = c.take(sizeof(uint2) * ne),
               os0 = c.take(n_in), os1 = c.take(n_out), ocd = c.take(sizeof(int) * n_out), opm = c.take(sizeof(int) * PM_WORDS), od = c.take(sizeof(ProblemDesc)),
               ocsd = c.take(ne * 32), oint = c.take(sizeof(int) * (ne + 1));
        buf.ensure(c.off, false);
        char *b = (char *)buf.p;
        CK(cudaMemcpyAsync(b + ok, kernel, sizeof(float) * ne, cudaMemcpyHostToDevice, g_stream));
        int nbits = 0;
        if (center) {
            std::vector<float> q(3 * (size_t)n_in, 1.0f); // non-zero ranges: no row is blanked
            CK(cudaMemcpyAsync(b + oq, q.data(), sizeof(float) * 3 * n_in, cudaMemcpyHostToDevice, g_stream));
            ProblemDesc d;
            memset(&d, 0, sizeof(d));
            d.n_in = (int)n_in;
            d.n_out = (int)n_out;
            d.kernel = (const float *)(b + ok);
            d.qint = (const float *)(b + oq);
            d.lat = (const float *)(b + ol);
            d.masks0 = (uint2 *)(b + om);
            d.shift0 = (int8_t *)(b + os0);
            d.shift1 = (int8_t *)(b + os1);
            d.col_digits = (int *)(b + ocd);
            d.prep_meta = (int *)(b + opm);
            CK(cudaMemcpyAsync(b + od, &d, sizeof(d), cudaMemcpyHostToDevice, g_stream));
            cmvm_prep_kernel<<<1, 256, 0, g_stream>>>((ProblemDesc *)(b + od));
            CK(cudaGetLastError());
            int pm[PM_WORDS];
            CK(cudaMemcpyAsync(pm, b + opm, sizeof(pm), cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            nbits = pm[PM_NBITS];
            csd_export_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, g_stream>>>((const uint2 *)(b + om), (int)ne, nbits, (int8_t *)(b + ocsd));
            CK(cudaMemcpyAsync(shift0, b + os0, n_in, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaMemcpyAsync(shift1, b + os1, n_out, cudaMemcpyDeviceToHost, g_stream));
        }
        else {
            int *xi = (int *)(b + oint);
            float_to_int_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, g_stream>>>((const float *)(b + ok), (int)ne, xi);
            CK(cudaMemsetAsync(xi + ne, 0, sizeof(int), g_stream));
            int_absmax_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, g_stream>>>(xi, (int)ne, xi + ne);
            int mx = 0;
            CK(cudaMemcpyAsync(&mx, xi + ne, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
            nbits = std::max(1, ceil_log2_pos((double)std::max((float)mx, 1.0f) * 1.5));
            int_csd_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, g_stream>>>(xi, (int)ne, nbits, (int8_t *)(b + ocsd));
            memset(shift0, 0, n_in);
            memset(shift1, 0, n_out);
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(csd, b + ocsd, ne * nbits, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        *n_bits = nbits;
    });
}

int da4ml_cmvm_int_arr_to_csd(const int32_t *x, int64_t n, int8_t *csd, int64_t *n_bits) {
    return guarded([&] {
        if (!x || n <= 0 || !csd || !n_bits)
            throw ApiError(DA4ML_E_INVALID, "invalid argument");
        init_device();
        static DevBuf buf;
        Carver c;
        size_t ox = c.take(sizeof(int) * (n + 1)), oc = c.take((size_t)n * 32);
        buf.ensure(c.off, false);
        char *b = (char *)buf.p;
        int *xi = (int *)(b + ox);
        CK(cudaMemcpyAsync(xi, x, sizeof(int) * n, cudaMemcpyHostToDevice, g_stream));
        CK(cudaMemsetAsync(xi + n, 0, sizeof(int), g_stream));
        int_absmax_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(xi, (int)n, xi + n);
        int mx = 0;
        CK(cudaMemcpyAsync(&mx, xi + n, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        int nbits = std::max(1, ceil_log2_pos((double)std::max((float)mx, 1.0f) * 1.5));
        int_csd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(xi, (int)n, nbits, (int8_t *)(b + oc));
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(csd, b + oc, (size_t)n * nbits, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        *n_bits = nbits;
    });
}

int da4ml_cmvm_kernel_decompose(const float *kernel, int64_t n_in, int64_t n_out, int dc, float *m0, float *m1) {
    return guarded([&] {
        if (!kernel || n_in <= 0 || n_out <= 0 || !m0 || !m1)
            throw ApiError(DA4ML_E_INVALID, "csd_decompose only supports 2D arrays.");
        init_device();
        const size_t n = (size_t)n_out + 1;
        static DevBuf buf;
        Carver c;
        size_t ok = c.take(sizeof(float) * n_in * n_out), oa = c.take(sizeof(float) * n_in * n), od = c.take(sizeof(int) * n * n), os = c.take(n * n), os0 = c.take(n_in),
               os1 = c.take(n_out), om0 = c.take(sizeof(float) * n_in * n_out), om1 = c.take(sizeof(float) * n_out * n_out), omap = c.take(sizeof(int) * 2 * n), oj = c.take(sizeof(DecompJob));
        buf.ensure(c.off, false);
        char *b = (char *)buf.p;
        CK(cudaMemcpyAsync(b + ok, kernel, sizeof(float) * n_in * n_out, cudaMemcpyHostToDevice, g_stream));
        center_kernel<<<1, 256, 0, g_stream>>>((const float *)(b + ok), (int)n_in, (int)n_out, (float *)(b + oa), (int8_t *)(b + os0), (int8_t *)(b + os1));
        long long tot = (long long)n * n;
        dist_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, g_stream>>>((const float *)(b + oa), (int)n_in, (int)n, (int *)(b + od), (int8_t *)(b + os));
        DecompJob j{dc, (float *)(b + om0), (float *)(b + om1), (int *)(b + omap)};
        CK(cudaMemcpyAsync(b + oj, &j, sizeof(j), cudaMemcpyHostToDevice, g_stream));
        const int threads = std::min(1024, std::max(64, (int)((n + 31) / 32 * 32)));
        const size_t smem = n * 17 + 64;
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(mst_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mst_build_kernel<<<1, threads, smem, g_stream>>>((const float *)(b + oa), (const int *)(b + od), (const int8_t *)(b + os), (const int8_t *)(b + os0), (const int8_t *)(b + os1), (int)n_in, (int)n_out, (const DecompJob *)(b + oj));
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(m0, b + om0, sizeof(float) * n_in * n_out, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaMemcpyAsync(m1, b + om1, sizeof(float) * n_out * n_out, cudaMemcpyDeviceToHost, g_stream));
        CK(cudaStreamSynchronize(g_stream));
    });
}

// developer probe: microseconds per group exchange at group size G (work = dummy stores per thread per round)
int da4ml_cmvm_debug_xchg_bench(int G, int iters, int work, double *us_per_iter) {
    return guarded([&] {
        init_device();
        static DevBuf buf;
        Carver c;
        size_t ob = c.take(256), ox = c.take(sizeof(unsigned long long) * 8 * G), os = c.take(sizeof(unsigned) * 512 * 8 * (size_t)G + 4096), oc = c.take(sizeof(long long) * G);
        buf.ensure(c.off, false);
        char *b = (char *)buf.p;
        CK(cudaMemsetAsync(b, 0, c.off, g_stream));
        GroupWs ws;
        memset(&ws, 0, sizeof(ws));
        ws.barrier = (unsigned *)(b + ob);
        ws.xchg = (unsigned long long *)(b + ox);
        unsigned *sink = (unsigned *)(b + os);
        long long *cyc = (long long *)(b + oc);
        void *args[] = {(void *)&ws, (void *)&G, (void *)&iters, (void *)&work, (void *)&sink, (void *)&cyc};
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, g_stream));
        CK(cudaLaunchCooperativeKernel((void *)xchg_bench_kernel, dim3(G), dim3(512), args, 0, g_stream));
        CK(cudaEventRecord(e1, g_stream));
        CK(cudaStreamSynchronize(g_stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *us_per_iter = 1e3 * ms / iters;
    });
}

// ---- DAIS program replay (reference dais/bindings.cc `run_interp`), opcodes -1/0/1 -------------------------------
int da4ml_dais_run(const int32_t *program, int64_t n_words, const double *inputs, int64_t n_samples, double *outputs) {
    return guarded([&] {
        if (!program || n_words < 6 || !inputs || !outputs || n_samples < 0)
            throw ApiError(DA4ML_E_RUNTIME, "Binary data too small to contain valid DAIS model file"); // DAISInterpreter.cc:11-15
        if (program[0] != 1)
            throw ApiError(DA4ML_E_RUNTIME, "DAIS version mismatch: expected version 1, got version " + std::to_string(program[0]));
        const int n_in = program[2], n_out = program[3], n_ops = program[4], n_tables = program[5];
        if (n_tables != 0 || n_words != 6 + (int64_t)n_in + 3LL * n_out + 8LL * n_ops)
            throw ApiError(DA4ML_E_RUNTIME, "Binary data size mismatch (lookup tables are not produced by the CMVM path)");
        if (n_samples == 0)
            return;
        init_device();
        const int32_t *inp_shifts = program + 6, *out_idxs = inp_shifts + n_in, *out_shifts = out_idxs + n_out, *out_negs = out_shifts + n_out;
        const DaisOp *hops = reinterpret_cast<const DaisOp *>(out_negs + n_out);
        // causality + supported opcodes (DAISInterpreter::validate), levels
        std::vector<int> level(n_ops, 0);
        int n_levels = 1;
        for (int i = 0; i < n_ops; ++i) {
            const DaisOp &op = hops[i];
            if (op.opcode == -1) {
                if (op.id0 < 0 || op.id0 >= n_in)
                    throw ApiError(DA4ML_E_RUNTIME, "input index out of range at operation " + std::to_string(i));
                continue;
            }
            if (op.opcode != 0 && op.opcode != 1)
                throw ApiError(DA4ML_E_RUNTIME, "Unknown opcode: " + std::to_string(op.opcode) + " at index " + std::to_string(i) + " (only adder graphs are replayed on this path)");
            if (op.id0 < 0 || op.id0 >= i || op.id1 < 0 || op.id1 >= i)
                throw ApiError(DA4ML_E_RUNTIME, "Operation " + std::to_string(i) + " violating causality");
            level[i] = 1 + std::max(level[op.id0], level[op.id1]);
            n_levels = std::max(n_levels, level[i] + 1);
        }
        std::vector<int> lvl_begin(n_levels + 1, 0), order(n_ops);
        for (int i = 0; i < n_ops; ++i)
            lvl_begin[level[i] + 1]++;
        for (int l = 0; l < n_levels; ++l)
            lvl_begin[l + 1] += lvl_begin[l];
        {
            std::vector<int> fill(lvl_begin.begin(), lvl_begin.end() - 1);
            for (int i = 0; i < n_ops; ++i)
                order[fill[level[i]]++] = i;
        }
        // samples per chunk: keep the [n_ops][S] int64 buffer around 1 GB
        const long long S_max = std::max<long long>(256, (1LL << 27) / std::max(1, n_ops));
        const long long S = std::min<long long>(n_samples, S_max);
        static DevBuf buf;
        Carver c;
        size_t o_ops = c.take(sizeof(DaisOp) * n_ops), o_ord = c.take(sizeof(int) * n_ops), o_hdr = c.take(sizeof(int) * (n_in + 3 * n_out)),
               o_in = c.take(sizeof(double) * S * n_in), o_out = c.take(sizeof(double) * S * n_out), o_buf = c.take(sizeof(long long) * S * n_ops);
        buf.ensure(c.off, false);
        char *b = (char *)buf.p;
        CK(cudaMemcpyAsync(b + o_ops, hops, sizeof(DaisOp) * n_ops, cudaMemcpyHostToDevice, g_stream));
        CK(cudaMemcpyAsync(b + o_ord, order.data(), sizeof(int) * n_ops, cudaMemcpyHostToDevice, g_stream));
        CK(cudaMemcpyAsync(b + o_hdr, inp_shifts, sizeof(int) * (n_in + 3 * n_out), cudaMemcpyHostToDevice, g_stream));
        const int *d_inp_shifts = (const int *)(b + o_hdr), *d_out_idxs = d_inp_shifts + n_in, *d_out_shifts = d_out_idxs + n_out, *d_out_negs = d_out_shifts + n_out;
        for (int64_t s0 = 0; s0 < n_samples; s0 += S) {
            const long long Sc = std::min<long long>(S, n_samples - s0);
            CK(cudaMemcpyAsync(b + o_in, inputs + s0 * n_in, sizeof(double) * Sc * n_in, cudaMemcpyHostToDevice, g_stream));
            for (int l = 0; l < n_levels; ++l) {
                const int n_l = lvl_begin[l + 1] - lvl_begin[l];
                if (!n_l)
                    continue;
                const long long threads = (long long)n_l * Sc;
                dais_level_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, g_stream>>>((const DaisOp *)(b + o_ops), (const int *)(b + o_ord) + lvl_begin[l], n_l, d_inp_shifts, (const double *)(b + o_in), n_in, Sc, (long long *)(b + o_buf));
            }
            const long long threads = (long long)n_out * Sc;
            dais_output_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, g_stream>>>((const DaisOp *)(b + o_ops), d_out_idxs, d_out_shifts, d_out_negs, n_out, Sc, (const long long *)(b + o_buf), (double *)(b + o_out));
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(outputs + s0 * n_out, b + o_out, sizeof(double) * Sc * n_out, cudaMemcpyDeviceToHost, g_stream));
            CK(cudaStreamSynchronize(g_stream));
        }
    });
}

int da4ml_cmvm_get_lsb_loc(float x) { return get_lsb_loc(x); }
int da4ml_cmvm_iceil_log2(float x) { return iceil_log2(x); }
int da4ml_cmvm_cost_add(const float q0[3], const float q1[3], int64_t shift, int sub, int adder_size, int carry_size, float out[2]) {
    cost_add(QInt{q0[0], q0[1], q0[2]}, QInt{q1[0], q1[1], q1[2]}, shift, sub != 0, adder_size, carry_size, out[0], out[1]);
    return DA4ML_OK;
}

} // extern "C"
