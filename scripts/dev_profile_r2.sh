# Round-2 profile captures (under gpurun, one GPU): launch list of one bench step, full-set capture of the dominant launch.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --c4 0 --cpu-seconds 0 > gpurun_out/launches_r2.log 2>&1
tail -2 gpurun_out/launches_r2.log | cut -c1-300
# the stage-0 launch of the default solve (6 shared jobs x 24 CTAs): second launch of cmvm_solve_kernel in the process after the warm-up solve
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:cmvm_solve_kernel -s 4 -c 1 -f -o gpurun_out/prof_r2_stage0 python scripts/dev_ncu_default.py > gpurun_out/ncu_r2.log 2>&1
tail -3 gpurun_out/ncu_r2.log
ls -la gpurun_out
