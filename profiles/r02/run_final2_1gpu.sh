#!/bin/bash
mkdir -p gpurun_out/final; O=gpurun_out/final
bash profiles/r02/run_bench_1gpu.sh
ncu --set full --clock-control none --import-source on -k regex:k_project_stream -s 3 -c 1 -o $O/prof_k1_c4s python bench.py --workload c4s --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/prof_k1_c4s.log 2>&1
ls $O | head -30
