mkdir -p gpurun_out/r2i
python bench.py --steps 20 --warmup 3 > gpurun_out/r2i/bench_n1.json 2> gpurun_out/r2i/bench_n1.err
tail -c 600 gpurun_out/r2i/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i/launches_bench_c4.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu --no-e2e --no-nodict > gpurun_out/r2i/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:heat_ -c 2 -f -o gpurun_out/r2i/prof_c4_r2 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 1 > gpurun_out/r2i/ncu_full.log 2>&1
ls -la gpurun_out/r2i
