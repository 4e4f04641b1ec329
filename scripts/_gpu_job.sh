ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ngm:: -c 4000 --csv --log-file gpurun_out/r1z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1z_ncu_bench.log 2>&1
grep -c ngm gpurun_out/r1z_launches.csv; tail -2 gpurun_out/r1z_ncu_bench.log | cut -c1-200
