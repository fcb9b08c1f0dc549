mkdir -p gpurun_out/r2z
for sec in 4 2 1; do
  timeout 400 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --dedup 0 --sector $sec --optimize 1 > gpurun_out/r2z/nodict_s$sec.log 2>&1
  echo "no dictionary, optimised, sector $sec: $(grep 'rep 3' gpurun_out/r2z/nodict_s$sec.log | cut -c1-60)"
done
