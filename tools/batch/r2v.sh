mkdir -p gpurun_out/r2v
python -m pytest tests -m gpu -x -q > gpurun_out/r2v/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2v/pytest_gpu.log
tail -n 4 gpurun_out/r2v/pytest_gpu.log
python bench.py > gpurun_out/r2v/bench_default.json 2> gpurun_out/r2v/bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/r2v/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_ifunction','ms_ijacobian','gpu_launches')}); print(d['e2e']); print(d['cpu_baseline']); print(d['roofline']['frac'], d['roofline']['traffic'], d['roofline'].get('ijacobian_no_dictionary'))
"
