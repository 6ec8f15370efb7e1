# development: run scripts/dev_variant_run.py with da4ml_b200/_binary/variants/$V.so in place of the library
cp da4ml_b200/_binary/libda4ml_b200_cmvm.so /tmp/keep.so
cp da4ml_b200/_binary/variants/$V.so da4ml_b200/_binary/libda4ml_b200_cmvm.so
timeout 300 python scripts/dev_variant_run.py $V ${G:-24} 2>&1 | tail -${TAIL:-40}
cp /tmp/keep.so da4ml_b200/_binary/libda4ml_b200_cmvm.so
