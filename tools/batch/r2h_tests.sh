mkdir -p gpurun_out/r2h
python -m pytest tests -m gpu -x -q > gpurun_out/r2h/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2h/pytest_gpu.log
tail -n 15 gpurun_out/r2h/pytest_gpu.log
