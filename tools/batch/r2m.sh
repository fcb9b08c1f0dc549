mkdir -p gpurun_out/r2m
python -m pytest tests/test_gpu_parity.py -x -q -k "cached_metric or overlapped or pattern_dictionary or gather" > gpurun_out/r2m/pytest.log 2>&1; tail -3 gpurun_out/r2m/pytest.log
for cm in 1 0; do
  python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --cache-metric $cm > gpurun_out/r2m/cm_$cm.log 2>&1
  echo "cache_metric $cm: $(grep 'rep 3' gpurun_out/r2m/cm_$cm.log | cut -c1-60) $(grep 'J sum' gpurun_out/r2m/cm_$cm.log)"
  python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --cache-metric $cm --dedup 0 > gpurun_out/r2m/cm_${cm}_nodict.log 2>&1
  echo "cache_metric $cm no dictionary: $(grep 'rep 3' gpurun_out/r2m/cm_${cm}_nodict.log | cut -c1-60)"
done
