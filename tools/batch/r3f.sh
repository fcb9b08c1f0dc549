mkdir -p gpurun_out/r3f
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "morton or lane_order or pattern_dictionary" > gpurun_out/r3f/pytest.log 2>&1; tail -3 gpurun_out/r3f/pytest.log
( time timeout 600 python tools/time_config.py -n 184 --tile 0 --numbering morton --block-rows 192 --scatter 2 --reps 4 ) > gpurun_out/r3f/morton_192.log 2>&1
echo "morton block_rows 192: $(grep 'rep 3' gpurun_out/r3f/morton_192.log | cut -c1-60); $(grep 'engine create' gpurun_out/r3f/morton_192.log | cut -c1-40); $(grep real gpurun_out/r3f/morton_192.log)"
