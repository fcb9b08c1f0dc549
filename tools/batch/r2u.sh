mkdir -p gpurun_out/r2u
timeout 600 ncu --set full --clock-control none --import-source on -k regex:residual_tiles -c 1 -f -o gpurun_out/r2u/prof_rt4 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 1 --rtiles 4 > gpurun_out/r2u/ncu.log 2>&1
ls -la gpurun_out/r2u
