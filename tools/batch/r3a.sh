mkdir -p gpurun_out/r3a
for th in 512 1024; do
  timeout 300 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --threads $th > gpurun_out/r3a/th_$th.log 2>&1
  echo "threads $th: $(grep 'rep 3' gpurun_out/r3a/th_$th.log | cut -c1-60)"
done
