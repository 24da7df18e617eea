#!/bin/bash
# Round-2 single-GPU evidence run (under gpurun): tests, the driver's two bench commands, every BASELINE config, launch
# lists and one full ncu capture per dominant kernel.  Outputs land in gpurun_out/final/.
mkdir -p gpurun_out/final; O=gpurun_out/final
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -1 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 300 $O/bench_c4.json
python bench.py --impl reference --steps 20 --warmup 5 > $O/reference_c4.json 2> $O/reference_c4.err; cut -c1-260 $O/reference_c4.json
for wl in c1 c2 c4s c5; do python bench.py --workload $wl --steps 20 --warmup 5 --no-e2e > $O/bench_$wl.json 2> $O/bench_$wl.err; cut -c1-160 $O/bench_$wl.json; done
python bench.py --workload c3 --steps 10 --warmup 3 --no-e2e --progressive > $O/bench_c3.json 2> $O/bench_c3.err; cut -c1-160 $O/bench_c3.json
for wl in c4 c2; do ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_project|k_tile|k_bin|k_queue|k_colormap" -s 7 -c 28 --csv --log-file $O/launches_$wl.csv python bench.py --workload $wl --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:k_project_stream -s 3 -c 1 -o $O/prof_k1_c4 python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/prof_k1_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_project_stream -s 3 -c 1 -o $O/prof_k1_c4s python bench.py --workload c4s --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/prof_k1_c4s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_gather -s 1 -c 1 -o $O/prof_k3_c2 python bench.py --workload c2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/prof_k3_c2.log 2>&1
ncu --set full --clock-control none -k regex:k_colormap -s 1 -c 1 -o $O/prof_k5_c4 python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/prof_k5_c4.log 2>&1
ls -la $O | head -40
