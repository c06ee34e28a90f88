#!/bin/bash
# final round-2 GPU call: full GPU suite, bench line, BB-144 lines, ncu evidence, reduced example grid
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_final_tests.log
tail -4 gpurun_out/r2_final_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; cut -c1-400 gpurun_out/r2_final_bench.json
python scripts/bench_bb144.py --p 0.003 > gpurun_out/r2_final_bb144_bp.json 2> gpurun_out/r2_final_bb144_bp.err; cut -c1-300 gpurun_out/r2_final_bb144_bp.json
python scripts/bench_bb144.py --p 0.003 --osd > gpurun_out/r2_final_bb144_bposd.json 2> gpurun_out/r2_final_bb144_bposd.err; cut -c1-300 gpurun_out/r2_final_bb144_bposd.json
python scripts/bench_bb144.py --p 0.001 --osd --no-cpu-baseline > gpurun_out/r2_final_bb144_bposd_p1e3.json 2>> gpurun_out/r2_final_bb144_bposd.err; cut -c1-200 gpurun_out/r2_final_bb144_bposd_p1e3.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:osd_rref_kernel -c 1 -o gpurun_out/r2_osd_rref_kernel -f python scripts/bp_bench.py 0.003 2368 30 osd 0 > gpurun_out/ncu_osd_rref.log 2>&1; tail -1 gpurun_out/ncu_osd_rref.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final_launches.log 2>&1
S=$(date +%s); timeout 170 python examples/rot_vs_unrot_surface_code.py --distances 9 11 13 17 --errors 0.002 0.004 0.006 0.008 --max-shots 300000 --out gpurun_out/r2_example > gpurun_out/r2_example.log 2> gpurun_out/r2_example.err; echo "rc $? wall seconds: $(( $(date +%s) - S ))" >> gpurun_out/r2_example.log; tail -3 gpurun_out/r2_example.log
