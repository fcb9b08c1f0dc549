mkdir -p gpurun_out/r2q
python -m pytest tests/test_gpu_parity.py tests/test_gpu_partition.py -x -q -k "host or pipelined" > gpurun_out/r2q/pytest.log 2>&1; tail -3 gpurun_out/r2q/pytest.log
python bench.py --steps 20 --warmup 3 --no-configs > gpurun_out/r2q/bench_n1.json 2> gpurun_out/r2q/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2q/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_ifunction','ms_ijacobian','gpu_launches')}); print(d['e2e'])
"
