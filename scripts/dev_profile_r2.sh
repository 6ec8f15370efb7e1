# Round-2 evidence (under gpurun, one GPU): GPU tests, smoke, bench line, launch list of one bench step, full-set capture of
# the dominant launch, SASS opcode histogram.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> gpurun_out/r2_gputests.txt
cat gpurun_out/r2_gputests.txt
timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/bench_r2_n1.err | tail -1 > gpurun_out/bench_r2_n1.json
cut -c1-400 gpurun_out/bench_r2_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> /dev/null | tail -1 > gpurun_out/bench_r2_reference.json
cut -c1-300 gpurun_out/bench_r2_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --c4 0 --cpu-seconds 0 > gpurun_out/launches_r2.log 2>&1
# the stage-0 launch of the default solve (6 shared jobs x 24 CTAs): launch order per solve = stage 0, stage 1 -> `-s 2 -c 1` is the second solve's stage 0
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:cmvm_solve_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_stage0 python scripts/dev_ncu_default.py > gpurun_out/ncu_r2.log 2>&1
tail -2 gpurun_out/ncu_r2.log
timeout 300 python scripts/dev_variant_run.py solve 24 2>&1 | tail -9 > gpurun_out/r2_phase_profile.txt
timeout 300 python scripts/dev_variant_run.py solve 148 2>&1 | tail -9 >> gpurun_out/r2_phase_profile.txt
cat gpurun_out/r2_phase_profile.txt
ls -la gpurun_out
