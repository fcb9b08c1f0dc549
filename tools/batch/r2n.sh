mkdir -p gpurun_out/r2n
python bench.py --steps 20 --warmup 3 > gpurun_out/r2n/bench_n1.json 2> gpurun_out/r2n/bench_n1.err
tail -c 400 gpurun_out/r2n/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'heat_|set_values|block_metrics' -c 60 --csv --log-file gpurun_out/r2n/launches_bench_c4.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu --no-e2e --no-nodict > gpurun_out/r2n/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:heat_ -c 2 -f -o gpurun_out/r2n/prof_c4_r2n python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 1 > gpurun_out/r2n/ncu_full.log 2>&1
ls -la gpurun_out/r2n
