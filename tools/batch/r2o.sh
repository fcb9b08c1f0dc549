mkdir -p gpurun_out/r2o
python -m pytest tests -m gpu -x -q > gpurun_out/r2o/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/r2o/pytest_gpu.log
tail -n 5 gpurun_out/r2o/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2o/bench_n1.json 2> gpurun_out/r2o/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2o/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_ifunction','ms_ijacobian')})
for k,v in d['configs'].items():
    if k!='fp64_peak': print(k, v.get('ms_ifunction'), v.get('ms_ijacobian'), v.get('roofline',{}).get('frac'), v.get('error'))
"
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2o/smoke.log 2>&1; tail -2 gpurun_out/r2o/smoke.log
