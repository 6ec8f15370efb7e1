# development A/B (under gpurun): the shipped library and the variants in da4ml_b200/_binary/variants/*.so on the bench stage (256x256, G = 24),
# the default solve and BASELINE config 4 at several group sizes; then the GPU suite with the shipped library
mkdir -p gpurun_out
cp da4ml_b200/_binary/libda4ml_b200_cmvm.so /tmp/keep.so
for v in ${VARIANTS:-shipped}; do
  if [ "$v" != shipped ]; then cp da4ml_b200/_binary/variants/$v.so da4ml_b200/_binary/libda4ml_b200_cmvm.so; else cp /tmp/keep.so da4ml_b200/_binary/libda4ml_b200_cmvm.so; fi
  echo "== $v"
  timeout 300 python scripts/dev_variant_run.py $v 24 2>&1 | tail -8
  GS=${GS:-0,3,4} timeout 300 python scripts/dev_c4_profile.py 2>&1 | grep -v "steps \.\."
done 2>&1 | tee gpurun_out/ab.txt
cp /tmp/keep.so da4ml_b200/_binary/libda4ml_b200_cmvm.so
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/ab_gputests.txt
