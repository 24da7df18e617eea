#!/bin/bash
# The bench half of run_final_1gpu.sh (no ncu): smoke, tests, the driver's two commands, every BASELINE config.
mkdir -p gpurun_out/final; O=gpurun_out/final
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -1 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 200 $O/bench_c4.json
python bench.py --impl reference --steps 20 --warmup 5 > $O/reference_c4.json 2> $O/reference_c4.err; cut -c1-200 $O/reference_c4.json
for wl in c1 c2 c4s c5; do python bench.py --workload $wl --steps 20 --warmup 5 --no-e2e > $O/bench_$wl.json 2> $O/bench_$wl.err; cut -c1-160 $O/bench_$wl.json; done
python bench.py --workload c3 --steps 10 --warmup 3 --no-e2e --progressive > $O/bench_c3.json 2> $O/bench_c3.err; cut -c1-160 $O/bench_c3.json
