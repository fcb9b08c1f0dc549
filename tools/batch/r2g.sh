mkdir -p gpurun_out/r2g
for sec in 4 2 1; do
  for th in 512 1024; do
    python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --dedup 0 --sector $sec --threads $th > gpurun_out/r2g/nodict_s${sec}_t$th.log 2>&1
    echo "dedup 0 sector $sec threads $th: $(grep 'rep 3' gpurun_out/r2g/nodict_s${sec}_t$th.log | awk '{print $6}') ms; $(grep '^plan' gpurun_out/r2g/nodict_s${sec}_t$th.log | cut -c1-150)"
  done
done
