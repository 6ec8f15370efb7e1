# Developer timing (under gpurun): default solve with and without sharing of identical jobs, a batch of 128x128 int6.
mkdir -p gpurun_out
python - <<PY > gpurun_out/r2_share.log 2>&1
import sys, time
sys.path.insert(0, ".")
import numpy as np
import da4ml_b200._binary as B
W = np.random.default_rng(0).integers(-128, 128, size=(256, 256)).astype(np.float32)
B.solve_raw(W[:8,:8].copy())
for share in (False, True):
    B.set_job_sharing(share)
    raw = B.solve_raw(W)
    t0 = time.time(); raw = B.solve_raw(W); t1 = time.time()
    print("share", share, "device ms", round(raw.device_ms, 1), "wall ms", round(1e3 * (t1 - t0), 1), "adders", raw.n_adders, "jobs", raw.profile["jobs_total"], raw.profile["jobs_run"], "G", raw.counters[0]["group_ctas"], "launches", raw.launches, flush=True)
W6 = [np.random.default_rng(s).integers(-32, 32, size=(128, 128)).astype(np.float32) for s in range(64)]
B.solve_batch_raw(W6[:16])
t0 = time.time(); rs = B.solve_batch_raw(W6[:16]); t1 = time.time()
print("batch16x128x6 wall ms", round(1e3 * (t1 - t0)), [r.n_adders for r in rs][:3], "jobs", rs[0].profile["jobs_total"], rs[0].profile["jobs_run"], "G", rs[0].counters[0]["group_ctas"], flush=True)
t0 = time.time(); rs = B.solve_batch_raw(W6); t1 = time.time()
print("batch64x128x6 wall ms", round(1e3 * (t1 - t0)), "jobs", rs[0].profile["jobs_total"], rs[0].profile["jobs_run"], "G", rs[0].counters[0]["group_ctas"], "device ms", round(rs[0].device_ms), "solve kernel ms", round(rs[0].profile["solve_kernel_ms"]), flush=True)
PY
cat gpurun_out/r2_share.log
