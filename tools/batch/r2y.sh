mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lane_order or pattern_dictionary or cached_metric" > gpurun_out/r2y/pytest.log 2>&1; tail -3 gpurun_out/r2y/pytest.log
for opt in 0 1; do
  timeout 400 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --dedup 0 --sector 2 --optimize $opt > gpurun_out/r2y/nodict_opt_$opt.log 2>&1
  echo "no dictionary, optimize $opt: $(grep 'rep 3' gpurun_out/r2y/nodict_opt_$opt.log | cut -c1-60); $(grep 'engine create' gpurun_out/r2y/nodict_opt_$opt.log | cut -c1-40)"
done
