mkdir -p gpurun_out/r2r
python tools/e2e_ab.py > gpurun_out/r2r/e2e_ab.log 2>&1; cat gpurun_out/r2r/e2e_ab.log | tail -6
