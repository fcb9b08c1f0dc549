mkdir -p gpurun_out/r3c
timeout 900 python -m pytest tests/test_examples_run.py -x -q > gpurun_out/r3c/pytest_examples.log 2>&1; tail -25 gpurun_out/r3c/pytest_examples.log
