// tsan_main.cc -- TEST INFRASTRUCTURE ONLY: race check of the kernel sources on the CPU.  Runs a few small solves (and
// the decomposition kernels) under the SIMT shim built with ThreadSanitizer; every simulated thread is a
// TSan fiber, barriers / collectives / acquire-release counters are the only synchronisation, so TSan reports conflicting
// accesses of the kernel code that nothing orders.
//   g++ -O1 -g -fsanitize=thread -DSIMT_TSAN -std=c++17 -ffp-contract=off -w tests/simt/tsan_main.cc -o /tmp/sim_tsan
//   setarch -R /tmp/sim_tsan        (TSan needs a fixed address-space layout on recent kernels)
#include "sim_cmvm.cc"

#include <random>

int main(int argc, char **argv) {
    if (argc > 1)
        sim_set_schedule(atoi(argv[1])); // thread schedule of the simulator (0 ascending, 1 descending, >= 2 pseudo-random)
    const int cases[][5] = {{8, 8, 4, 2, 64}, {12, 10, 6, 3, 64}, {6, 40, 5, 2, 64}, {16, 16, 6, 2, 64}}; // n_in, n_out, bits, ctas, threads
    int bad = 0;
    for (int variant = 0; variant < 2; ++variant) // planner's layout; then owner lists that spill to global memory and a small counter table
        for (auto &c : cases) {
            std::mt19937 rng(c[0] * 131 + c[1]);
            std::uniform_int_distribution<int> d(-(1 << (c[2] - 1)), (1 << (c[2] - 1)) - 1);
            std::vector<float> W((size_t)c[0] * c[1]), q, l(c[0], 0.0f);
            for (auto &v : W)
                v = (float)d(rng);
            for (int i = 0; i < c[0]; ++i)
                q.insert(q.end(), {-128.0f, 127.0f, 1.0f});
            const long long room = c[0] + (long long)c[0] * c[1] * 34 + 8;
            std::vector<int64_t> meta(32), is(c[0]), oi(c[1]), os(c[1]), on(c[1]), ops_i(4 * room);
            std::vector<float> ops_f(5 * room);
            if (c[0] == 16) // small segment: the compaction paths too
                sim_set_segment_cap(1500);
            if (variant == 1) {
                sim_set_list_cap(4);
                sim_set_own_caps(8, -1);
            }
            const long long n = sim_solve_single(W.data(), c[0], c[1], "wmc", q.data(), l.data(), -1, -1, c[3], c[4], 0, 2, meta.data(), is.data(), oi.data(), os.data(),
                                                 on.data(), ops_i.data(), ops_f.data(), room);
            sim_set_segment_cap(0);
            sim_set_list_cap(-1);
            sim_set_own_caps(0, -1);
            printf("%s %dx%d: %lld ops, %lld steps, %lld compactions%s\n", variant ? "spills " : "planned", c[0], c[1], n, (long long)meta[2], (long long)meta[9], n < 0 ? "  FAILED" : "");
            bad += n < 0;
            if (n == -100)
                printf("  %s\n", sim_last_error());
        }
    // accounting mode, wide (12-byte) list rows
    for (int mode = 0; mode < 2; ++mode) {
        const int n_in = 10, n_out = 12;
        std::mt19937 rng(77);
        std::vector<float> W((size_t)n_in * n_out), q, l(n_in, 0.0f);
        for (auto &v : W)
            v = (float)((int)(rng() % 63) - 31);
        for (int i = 0; i < n_in; ++i)
            q.insert(q.end(), {-128.0f, 127.0f, 1.0f});
        const long long room = n_in + (long long)n_in * n_out * 34 + 8;
        std::vector<int64_t> meta(32), is(n_in), oi(n_out), os(n_out), on(n_out), ops_i(4 * room);
        std::vector<float> ops_f(5 * room);
        sim_set_wide_rows(mode == 1);
        const long long n = sim_solve_single(W.data(), n_in, n_out, "wmc-dc", q.data(), l.data(), 2, 4, 3, 64, mode == 0, 2, meta.data(), is.data(), oi.data(), os.data(), on.data(),
                                             ops_i.data(), ops_f.data(), room);
        sim_set_wide_rows(0);
        printf("%s: %lld ops%s\n", mode == 0 ? "accounting" : "wide rows", n, n < 0 ? "  FAILED" : "");
        bad += n < 0;
    }
    // batched launch: three jobs on two groups (a group's workspace is reused by its second job)
    {
        const int n = 3, n_in[3] = {9, 6, 7}, n_out[3] = {8, 12, 7};
        std::vector<std::vector<float>> W(n), q(n), l(n), of(n);
        std::vector<std::vector<int64_t>> meta(n), is(n), oi(n), os(n), on(n), ops(n);
        std::vector<const float *> Wp(n), qp(n), lp(n);
        std::vector<float *> ofp(n);
        std::vector<int64_t *> mp(n), isp(n), oip(n), osp(n), onp(n), opp(n);
        std::vector<long long> room(n), got(n);
        std::mt19937 rng(91);
        for (int i = 0; i < n; ++i) {
            W[i].resize((size_t)n_in[i] * n_out[i]);
            for (auto &v : W[i])
                v = (float)((int)(rng() % 31) - 15);
            for (int k = 0; k < n_in[i]; ++k)
                q[i].insert(q[i].end(), {-128.0f, 127.0f, 1.0f});
            l[i].assign(n_in[i], 0.0f);
            room[i] = n_in[i] + (long long)n_in[i] * n_out[i] * 34 + 8;
            meta[i].resize(32), is[i].resize(n_in[i]), oi[i].resize(n_out[i]), os[i].resize(n_out[i]), on[i].resize(n_out[i]), ops[i].resize(4 * room[i]), of[i].resize(5 * room[i]);
            Wp[i] = W[i].data(), qp[i] = q[i].data(), lp[i] = l[i].data(), ofp[i] = of[i].data();
            mp[i] = meta[i].data(), isp[i] = is[i].data(), oip[i] = oi[i].data(), osp[i] = os[i].data(), onp[i] = on[i].data(), opp[i] = ops[i].data();
        }
        const int rc = sim_solve_many(n, Wp.data(), n_in, n_out, "wmc", qp.data(), lp.data(), 2, 2, 64, mp.data(), isp.data(), oip.data(), osp.data(), onp.data(), opp.data(), ofp.data(),
                                      room.data(), got.data());
        printf("batch: rc %d, ops %lld %lld %lld\n", rc, got[0], got[1], got[2]);
        bad += rc != 0 || got[0] < 0 || got[1] < 0 || got[2] < 0;
    }
    {
        std::vector<float> W(12 * 20), m0(12 * 20), m1(20 * 20);
        std::mt19937 rng(5);
        for (auto &v : W)
            v = (float)((int)(rng() % 255) - 127);
        bad += sim_kernel_decompose(W.data(), 12, 20, 1, m0.data(), m1.data()) != 0;
        printf("decompose 12x20 done\n");
    }
    return bad;
}
