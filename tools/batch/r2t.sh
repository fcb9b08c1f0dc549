mkdir -p gpurun_out/r2t
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "residual_tile" > gpurun_out/r2t/pytest_tiles.log 2>&1; tail -5 gpurun_out/r2t/pytest_tiles.log
for rt in 0 3 4; do
  timeout 200 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --rtiles $rt > gpurun_out/r2t/rt_$rt.log 2>&1
  echo "rtiles $rt: $(grep 'rep 3' gpurun_out/r2t/rt_$rt.log | cut -c1-60) $(grep 'F norm' gpurun_out/r2t/rt_$rt.log)"
done
