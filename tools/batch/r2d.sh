set -x
mkdir -p gpurun_out/r2d
python -m pytest tests/test_gpu_parity.py -k quadrilateral -x -q > gpurun_out/r2d/pytest_quad.log 2>&1; echo rc=$? >> gpurun_out/r2d/pytest_quad.log
python -m pytest tests/test_gpu_api.py -k "quadrilateral or cahn or split" -x -q > gpurun_out/r2d/pytest_quad_api.log 2>&1; echo rc=$? >> gpurun_out/r2d/pytest_quad_api.log
python -m pytest tests/test_adapter_fake.py -x -q > gpurun_out/r2d/pytest_adapter.log 2>&1; echo rc=$? >> gpurun_out/r2d/pytest_adapter.log
export FDTS_B200_LIB=$PWD/firedrake_ts_b200/csrc/libfdts_b200_dev.so
for dbg in 0 16 32 64 96 112 128 144 240; do
  python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 1 --reps 4 --debug $dbg > gpurun_out/r2d/resid_dbg_$dbg.log 2>&1
done
tail -n 4 gpurun_out/r2d/*.log
