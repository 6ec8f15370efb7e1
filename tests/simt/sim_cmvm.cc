// sim_cmvm.cc -- TEST INFRASTRUCTURE ONLY: runs the product's CUDA kernel sources (cmvm_prep_kernel, cmvm_solve_kernel;
// da4ml_b200/csrc/*.cuh, unmodified) on the CPU under the SIMT shim of simt.h, for one solve_single job, so that their
// logic can be compared with the oracle in the CPU test suite.  The buffer sizing mirrors run_stage_jobs
// (da4ml_b200/csrc/host_stage.cuh) and uses the product's own planner (host_plan.cuh).  Never linked into the product.
#include "simt.h"

#include "../../da4ml_b200/csrc/cmvm_kernels.cuh"
#include "../../da4ml_b200/csrc/cmvm_kernel_em.cuh"
#include "../../da4ml_b200/csrc/host_plan.cuh"

#include <string>

using namespace da;

namespace {
std::string g_err;
int method_id(const std::string &m) {
    const char *names[] = {"mc", "mc-dc", "mc-pdc", "wmc", "wmc-dc", "wmc-pdc", "dummy"};
    for (int i = 0; i < 7; ++i)
        if (m == names[i])
            return i;
    throw std::runtime_error("Unknown method: " + m);
}
int ilog2_ceil(int v) {
    int l = 0;
    while ((1 << l) < v)
        ++l;
    return l;
}
template <class T> T *zalloc(std::vector<std::unique_ptr<unsigned char[]>> &keep, size_t n) {
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T) + 64;
    keep.emplace_back(new unsigned char[bytes]());
    uintptr_t a = ((uintptr_t)keep.back().get() + 63) & ~(uintptr_t)63;
    return (T *)a;
}
} // namespace

extern "C" {

const char *sim_last_error() { return g_err.c_str(); }

// Self-test of the shim: warp collectives, block barrier, shared variables, inter-CTA polling, deadlock detection.
// Returns 0 when every check passes, else the number of the failing check.
int sim_selftest() {
    try {
        std::vector<int> out(2 * 64, -1);
        std::vector<unsigned> flag(1, 0u);
        simt::launch(dim3(2), dim3(64), 128, [&] {
            DA_SHARED_VAR(int, total);
            DA_DYN_SHARED(dyn);
            const int tid = threadIdx.x, lane = tid & 31;
            if (tid == 0)
                total = 0;
            __syncthreads();
            int v = lane;
            for (int off = 16; off > 0; off >>= 1)
                v += __shfl_xor_sync(0xffffffffu, v, off); // 496 on every lane
            const unsigned b = __ballot_sync(0xffffffffu, (lane & 1) != 0);
            const int up = __shfl_up_sync(0xffffffffu, lane, 1);
            if (v == 496 && b == 0xaaaaaaaau && up == (lane ? lane - 1 : 0) && __shfl_sync(0xffffffffu, tid, 5) == (tid & ~31) + 5)
                atomicAdd(&total, 1);
            ((int *)dyn)[tid & 31] = tid; // touch dynamic shared memory
            __syncthreads();
            // CTA 1 waits for CTA 0 through a polled global flag
            if (blockIdx.x == 0 && tid == 0)
                flag[0] = 1u;
            if (blockIdx.x == 1 && tid == 0)
                while (ld_acquire_u32(&flag[0]) == 0u) {
                }
            __syncthreads();
            out[blockIdx.x * 64 + tid] = total;
        });
        for (int v : out)
            if (v != 64)
                return 1;
        bool caught = false;
        try {
            simt::launch(dim3(1), dim3(32), 0, [&] {
                if (threadIdx.x != 3)
                    __syncthreads(); // thread 3 never arrives
            });
        } catch (const std::runtime_error &e) {
            caught = std::string(e.what()).find("deadlock") != std::string::npos;
        }
        if (!caught)
            return 2;
        caught = false;
        try {
            simt::launch(dim3(1), dim3(32), 0, [&] {
                if (threadIdx.x < 16)
                    (void)__ballot_sync(0xffffffffu, true); // half of the warp skips a full-mask collective
            });
        } catch (const std::runtime_error &) {
            caught = true;
        }
        return caught ? 0 : 3;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 100;
    }
}

// One solve_single on `G` simulated CTAs of `cta_threads` threads.  Returns the number of ops (>= 0) or -(status) when a
// capacity was exceeded, -100 on an exception (sim_last_error()).  meta_out: the kernel's 32 result words.
// `em` != 0: the expression-major kernel (cmvm_solve_em_kernel) instead of cmvm_solve_kernel.
long long sim_solve_single(const float *kernel, int n_in, int n_out, const char *method, const float *qint, const float *lat, int adder_size, int carry_size,
                           int G, int cta_threads, int global_lists, int accounting, int list_mul, int em, int64_t *meta_out, int64_t *inp_shifts, int64_t *out_idxs,
                           int64_t *out_shifts, int64_t *out_negs, int64_t *ops_i, float *ops_f, long long ops_room) {
    try {
        std::vector<std::unique_ptr<unsigned char[]>> keep;
        ProblemDesc d;
        memset(&d, 0, sizeof(d));
        d.n_in = n_in;
        d.n_out = n_out;
        d.method = method_id(method);
        d.adder_size = adder_size;
        d.carry_size = carry_size;
        float *k = zalloc<float>(keep, (size_t)n_in * n_out);
        memcpy(k, kernel, sizeof(float) * (size_t)n_in * n_out);
        d.kernel = k;
        float *q = zalloc<float>(keep, 3 * (size_t)n_in), *l = zalloc<float>(keep, n_in);
        memcpy(q, qint, sizeof(float) * 3 * n_in);
        memcpy(l, lat, sizeof(float) * n_in);
        d.qint = q;
        d.lat = l;
        d.masks0 = zalloc<uint2>(keep, (size_t)n_in * n_out);
        d.shift0 = zalloc<int8_t>(keep, n_in);
        d.shift1 = zalloc<int8_t>(keep, n_out);
        d.col_digits = zalloc<int>(keep, n_out);
        d.prep_meta = zalloc<int>(keep, PM_WORDS);
        simt::launch(dim3(1), dim3(256), 0, [&] { cmvm_prep_kernel(&d); });
        const int *pm = d.prep_meta;
        // ---- capacities (run_stage_jobs)
        const long long d0 = pm[PM_D0];
        d.nbits = pm[PM_NBITS];
        d.log_s = std::max(1, ilog2_ceil(2 * (2 * d.nbits - 1)));
        const long long t_cap = std::min<long long>(d0, d0 / 2 + 1024);
        d.e_cap = (int)(n_in + t_cap + 1);
        d.ops_cap = (int)(n_in + d0 + 1);
        d.col_cap = pm[PM_COLCAP] + 1;
        d.heap_lane_cap = (std::min(d.nbits, 32) + 1) * ((d.col_cap + 31) / 32) + 2;
        d.op_misc = zalloc<int4>(keep, d.ops_cap);
        d.op_q = zalloc<float4>(keep, d.ops_cap);
        d.op_cost = zalloc<float>(keep, d.ops_cap);
        d.out_q = zalloc<float4>(keep, n_out);
        d.cost_init = 0.0f;
        d.out_idx = zalloc<int>(keep, n_out);
        d.out_shift = zalloc<int>(keep, n_out);
        d.out_neg = zalloc<int>(keep, n_out);
        d.result_meta = zalloc<long long>(keep, META_WORDS);
        d.trace = nullptr;
        d.trace_cap = 0;
        // ---- launch geometry from the product's planner
        std::vector<PlanJob> pj(1);
        pj[0].n_in = n_in;
        pj[0].n_out = n_out;
        pj[0].nbits = d.nbits;
        pj[0].d0 = d0;
        pj[0].dcol_max = pm[PM_DCOL_MAX];
        pj[0].col_cap = d.col_cap;
        pj[0].list_mul = list_mul > 0 ? list_mul : 2;
        pj[0].global_lists = global_lists != 0 || em != 0;
        PlanEnv env;
        env.coop = G;
        env.group_override = G;
        env.accounting = accounting != 0;
        const LaunchPlan plan = plan_launch(pj, env);
        const LaunchCfg cfg = plan.cfg;
        // ---- group workspace
        GroupWs w;
        memset(&w, 0, sizeof(w));
        const long long max_heap = (long long)n_out * 32 * d.heap_lane_cap;
        w.col_u32 = zalloc<uint32_t>(keep, cfg.lcap > 0 ? 64 : (size_t)3 * n_out * d.col_cap);
        w.col_len = zalloc<int>(keep, n_out);
        w.col_k = zalloc<int>(keep, n_out);
        w.slab = zalloc<uint32_t>(keep, (size_t)3 * d.e_cap << d.log_s);
        w.mod_step = zalloc<uint32_t>(keep, d.e_cap);
        w.fseg = zalloc<FEnt>(keep, (size_t)G * plan.max_fcap);
        w.touch = zalloc<uint32_t>(keep, (size_t)G * plan.max_touch);
        w.slots = zalloc<uint4>(keep, 2 * (size_t)G);
        w.heap = zalloc<uint4>(keep, 2 * (size_t)max_heap);
        w.barrier = zalloc<unsigned>(keep, 64);
        w.xchg = zalloc<unsigned long long>(keep, 2 * 4 * (size_t)G);
        w.fseg_cap = (int)plan.max_fcap;
        w.touch_cap = (int)plan.max_touch;
        w.heap_cap = max_heap;
        if (!em)
            simt::launch(dim3(G), dim3(cta_threads), plan.smem_bytes, [&] { cmvm_solve_kernel(&d, 1, &w, cfg); });
        else {
            EmWs e;
            memset(&e, 0, sizeof(e));
            e.pool_cap = (int)(((long long)n_in * n_out + d0) / G + n_out + 64);
            e.words = (n_out + 31) / 32;
            e.e_cap = d.e_cap;
            e.cell_col = zalloc<uint32_t>(keep, (size_t)G * e.pool_cap);
            e.cell_pl[0] = zalloc<uint2>(keep, (size_t)G * e.pool_cap);
            e.cell_pl[1] = zalloc<uint2>(keep, (size_t)G * e.pool_cap);
            e.cell_off = zalloc<uint32_t>(keep, d.e_cap);
            e.cell_cnt = zalloc<uint32_t>(keep, d.e_cap);
            e.rowbits = zalloc<uint32_t>(keep, (size_t)d.e_cap * e.words);
            e.ver = zalloc<unsigned char>(keep, (size_t)G * d.e_cap);
            simt::launch(dim3(G), dim3(cta_threads), em_smem_bytes(cfg.nchunk_cap, n_out, cta_threads), [&] { cmvm_solve_em_kernel(&d, 1, &w, &e, cfg, n_out); });
        }
        for (int i = 0; i < META_WORDS; ++i)
            meta_out[i] = d.result_meta[i];
        meta_out[10] = d0;
        meta_out[11] = d.nbits;
        meta_out[12] = cfg.G;
        meta_out[15] = cfg.lcap;
        // the counter slab must be left zero for the next problem of the group
        for (size_t i = 0; i < ((size_t)3 * d.e_cap << d.log_s); ++i)
            if (w.slab[i] != 0u && d.result_meta[META_STATUS] == ST_OK)
                throw std::runtime_error("counter slab not left zero");
        if (d.result_meta[META_STATUS] != ST_OK)
            return -(long long)d.result_meta[META_STATUS];
        const long long n_ops = d.result_meta[META_N_OPS];
        if (n_ops > ops_room)
            throw std::runtime_error("ops_room too small");
        for (int i = 0; i < n_in; ++i)
            inp_shifts[i] = d.shift0[i];
        for (int o = 0; o < n_out; ++o) {
            out_idxs[o] = d.out_idx[o];
            out_shifts[o] = d.out_shift[o];
            out_negs[o] = d.out_neg[o];
        }
        for (long long i = 0; i < n_ops; ++i) {
            ops_i[4 * i + 0] = d.op_misc[i].x;
            ops_i[4 * i + 1] = d.op_misc[i].y;
            ops_i[4 * i + 2] = d.op_misc[i].z;
            ops_i[4 * i + 3] = d.op_misc[i].w;
            ops_f[5 * i + 0] = d.op_q[i].x;
            ops_f[5 * i + 1] = d.op_q[i].y;
            ops_f[5 * i + 2] = d.op_q[i].z;
            ops_f[5 * i + 3] = d.op_q[i].w;
            ops_f[5 * i + 4] = d.op_cost[i];
        }
        return n_ops;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -100;
    }
}
}
