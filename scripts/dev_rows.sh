# First GPU session with the expression-major kernel (DA4ML_B200_ROWS=1): parity, sanitizer, then A/B timing against the
# shipped kernel.  usage (on the GPU box): bash scripts/dev_rows.sh > gpurun_out/rows.log 2>&1
set -x
# 1. parity against the checker (opt-in test) -- stop here if it fails
DA4ML_B200_TEST_ROWS=1 timeout 600 python -m pytest tests/test_cmvm_gpu.py -x -q -m gpu -k rows_kernel 2>&1 | tail -5
# 2. races / barriers on a small case
DA4ML_B200_ROWS=1 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_case.py 2>&1 | tail -5
DA4ML_B200_ROWS=1 timeout 900 compute-sanitizer --tool synccheck python scripts/sanitize_case.py 2>&1 | tail -3
DA4ML_B200_ROWS=1 timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_case.py 2>&1 | tail -3
# 3. same library, both kernels: one 256x256 int8 stage at 14 CTAs, the full default solve, a batch (digest must agree)
timeout 300 python scripts/dev_variant_run.py columns 2>&1 | tail -1
DA4ML_B200_ROWS=1 timeout 300 python scripts/dev_variant_run.py rows 500 2000 6000 12000 2>&1 | tail -6
