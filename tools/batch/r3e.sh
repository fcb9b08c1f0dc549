mkdir -p gpurun_out/r3e
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "morton" > gpurun_out/r3e/pytest.log 2>&1; tail -4 gpurun_out/r3e/pytest.log
for br in 128 160 192; do
  timeout 600 python tools/time_config.py -n 184 --tile 0 --numbering morton --block-rows $br --scatter 2 --reps 4 > gpurun_out/r3e/morton_$br.log 2>&1
  echo "morton block_rows $br: $(grep 'rep 3' gpurun_out/r3e/morton_$br.log | cut -c1-60); $(grep '^plan' gpurun_out/r3e/morton_$br.log | cut -c1-140); $(tail -1 gpurun_out/r3e/morton_$br.log | cut -c1-200)"
done
