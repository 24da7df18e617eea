for lib in "" profiles/microbench/libtsplat_nostats.so; do
  export TSPLAT_LIBRARY=$lib; [ -z "$lib" ] && unset TSPLAT_LIBRARY
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l.csv python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
  echo "lib=$lib"; python profiles/launch_summary.py gpurun_out/l.csv | grep project_splat
done
