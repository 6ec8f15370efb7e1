# Round-2 (second session) evidence under gpurun, one GPU: GPU tests, smoke, bench line, reference arm, launch list of one bench step,
# compute-sanitizer on small solves, phase timers, config-4 profile.  (The device code is byte-identical to the build the
# ncu --set full capture of profiles/ncu_r2_* was taken from; see profiles/README.md.)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r2b_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> gpurun_out/r2b_gputests.txt
cat gpurun_out/r2b_gputests.txt
timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_r2b_n1.err | tail -1 > gpurun_out/bench_r2b_n1.json
cut -c1-600 gpurun_out/bench_r2b_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> /dev/null | tail -1 > gpurun_out/bench_r2b_reference.json
cut -c1-300 gpurun_out/bench_r2b_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 1 --warmup 1 --c4 0 --cpu-seconds 0 > gpurun_out/launches_r2b.log 2>&1
{ for tool in memcheck synccheck racecheck; do echo "== $tool"; timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_case.py 2>&1 | grep -i "summary\|hazard\|error" | tail -4; done; } > gpurun_out/sanitizer_r2.txt 2>&1
cat gpurun_out/sanitizer_r2.txt
timeout 300 python scripts/dev_variant_run.py solve 24 2>&1 | tail -9 > gpurun_out/phase_profile_r2b.txt
cat gpurun_out/phase_profile_r2b.txt
GS=0 timeout 300 python scripts/dev_c4_profile.py 2>&1 | tee gpurun_out/c4_profile_r2b.txt
ls -la gpurun_out
