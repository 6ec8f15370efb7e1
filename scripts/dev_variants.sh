# usage (on the GPU box): VARIANTS="a b" PROFILE_VARIANT=a bash scripts/dev_variants.sh   -- expects da4ml_b200/_binary/variants/<name>.so
set -x
timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 2 --warmup 3 --cpu-seconds 0 2>/dev/null | tail -1 > gpurun_out/bench_line.json; cut -c1-330 gpurun_out/bench_line.json
cp da4ml_b200/_binary/libda4ml_b200_cmvm.so /tmp/keep.so
for v in $VARIANTS; do cp da4ml_b200/_binary/variants/$v.so da4ml_b200/_binary/libda4ml_b200_cmvm.so; extra=""; if [ "$v" = "$PROFILE_VARIANT" ]; then extra="500 2000 6000 12000"; fi; timeout 120 python scripts/dev_variant_run.py $v $extra 2>&1 | tail -6; done
cp /tmp/keep.so da4ml_b200/_binary/libda4ml_b200_cmvm.so
