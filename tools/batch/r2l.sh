mkdir -p gpurun_out/r2l
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'heat_|set_values' -c 60 --csv --log-file gpurun_out/r2l/launches_bench_c4.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu --no-e2e --no-nodict > gpurun_out/r2l/ncu_launches.log 2>&1
grep -c heat_ gpurun_out/r2l/launches_bench_c4.csv
