mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_cmvm_gpu.py -x -q -m gpu -k "group_sizes or golden_full or random_options or dense_stack or job_sharing" 2>&1 | tail -4
python - <<PY
import sys, time
sys.path.insert(0, ".")
import numpy as np
import da4ml_b200._binary as B
W = np.random.default_rng(0).integers(-128, 128, size=(256, 256)).astype(np.float32)
B.solve_raw(W[:8,:8].copy())
for kind in ("columns", "owned"):
        for share in (False, True):
        B.set_job_sharing(share)
        raw = B.solve_raw(W)
        t0 = time.time(); raw = B.solve_raw(W); t1 = time.time()
        print(kind, "share", share, "device ms", round(raw.device_ms, 1), "wall ms", round(1e3 * (t1 - t0), 1), "adders", raw.n_adders, "jobs", raw.profile["jobs_total"], raw.profile["jobs_run"], "G", raw.counters[0]["group_ctas"], "launches", raw.launches, flush=True)
    W6 = [np.random.default_rng(s).integers(-32, 32, size=(128, 128)).astype(np.float32) for s in range(16)]
    t0 = time.time(); rs = B.solve_batch_raw(W6); t1 = time.time()
    print(kind, "batch16x128x6 wall ms", round(1e3 * (t1 - t0)), [r.n_adders for r in rs][:3], "jobs", rs[0].profile["jobs_total"], rs[0].profile["jobs_run"], "G", rs[0].counters[0]["group_ctas"], flush=True)
PY
) > gpurun_out/r2_share.log 2>&1
cat gpurun_out/r2_share.log
