mkdir -p gpurun_out/r3g
if [ "$1" = morton ]; then
timeout 900 ncu --set full --clock-control none -k regex:heat_jacobian_gather2 -c 1 -f -o gpurun_out/r3g/prof_morton python tools/time_config.py -n 184 --tile 0 --numbering morton --block-rows 192 --scatter 2 --reps 1 > gpurun_out/r3g/ncu_morton.log 2>&1
else
timeout 900 ncu --set full --clock-control none -k regex:heat_jacobian_gather2 -c 1 -f -o gpurun_out/r3g/prof_nodict python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 1 --dedup 0 > gpurun_out/r3g/ncu_nodict.log 2>&1
fi
ls -la gpurun_out/r3g
