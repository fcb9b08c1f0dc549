mkdir -p gpurun_out/r3i
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "untiled or morton or engine_matches" > gpurun_out/r3i/pytest.log 2>&1; tail -12 gpurun_out/r3i/pytest.log
