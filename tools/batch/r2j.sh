mkdir -p gpurun_out/r2j
python -m pytest tests/test_gpu_parity.py -k "residual_tile or tiles" -x -q > gpurun_out/r2j/pytest_tiles.log 2>&1; tail -3 gpurun_out/r2j/pytest_tiles.log
for rt in 0 2 3; do
  python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --rtiles $rt > gpurun_out/r2j/rt_$rt.log 2>&1
  echo "rtiles $rt: $(grep 'rep 3' gpurun_out/r2j/rt_$rt.log | cut -c1-60) $(grep 'F norm' gpurun_out/r2j/rt_$rt.log)"
done
