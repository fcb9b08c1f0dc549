mkdir -p gpurun_out/r2x
for opt in 0 2 1; do
  timeout 300 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 4 --optimize $opt > gpurun_out/r2x/opt_$opt.log 2>&1
  echo "optimize $opt: $(grep 'rep 3' gpurun_out/r2x/opt_$opt.log | cut -c1-60)"
done
