mkdir -p gpurun_out/r3b
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r3b/pytest.log 2>&1; tail -3 gpurun_out/r3b/pytest.log
for cm in 1 0; do
  timeout 300 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --cache-metric $cm > gpurun_out/r3b/cm_$cm.log 2>&1
  echo "cache_metric $cm: $(grep 'rep 3' gpurun_out/r3b/cm_$cm.log | cut -c1-60) $(grep 'J sum' gpurun_out/r3b/cm_$cm.log)"
done
timeout 300 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --dedup 0 > gpurun_out/r3b/nodict.log 2>&1
echo "no dictionary: $(grep 'rep 3' gpurun_out/r3b/nodict.log | cut -c1-60)"
