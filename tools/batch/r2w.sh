mkdir -p gpurun_out/r2w
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bbm_jacobian_tensor|residual_kernel' -c 2 -f -o gpurun_out/r2w/prof_c5 python tools/time_config.py --family bbm --dim 3 --degree 3 -n 48 --tile 2 --scatter 0 --reps 1 > gpurun_out/r2w/ncu_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jacobian_kernel|residual_kernel' -c 2 -f -o gpurun_out/r2w/prof_c3 python tools/time_config.py --family cahn_hilliard --dim 2 --degree 1 -n 2048 --tile 16 --scatter 0 --reps 1 > gpurun_out/r2w/ncu_c3.log 2>&1
ls -la gpurun_out/r2w; tail -3 gpurun_out/r2w/ncu_c5.log gpurun_out/r2w/ncu_c3.log
