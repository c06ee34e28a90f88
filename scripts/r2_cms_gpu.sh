#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bp or bposd or osd or empty_single or bb144" > gpurun_out/r2_cms_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_cms_tests.log
tail -5 gpurun_out/r2_cms_tests.log
{
for p in 0.003 0.002; do python scripts/bp_bench.py $p 8192 30 osd 0 2>&1 | tail -1; done
GQ_OSD_CHAIN=1 python scripts/bp_bench.py 0.003 8192 30 osd 0 2>&1 | tail -1 | sed "s/^/echelon walk: /"
python scripts/bp_bench.py 0.003 16384 30 osd 0 2>&1 | tail -1
} > gpurun_out/r2_cms_bench.log 2>&1
cat gpurun_out/r2_cms_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_bposd_launches.csv python scripts/bp_bench.py 0.003 8192 30 osd 0 > gpurun_out/ncu_bposd.log 2>&1
grep -E "bp_cms|osd_" gpurun_out/r2_bposd_launches.csv | tail -4
