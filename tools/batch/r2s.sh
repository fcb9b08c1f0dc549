mkdir -p gpurun_out/r2s
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2s/ref.json 2> gpurun_out/r2s/ref.err; tail -c 900 gpurun_out/r2s/ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
