mkdir -p gpurun_out/r3h
for opt in 1 3; do
  timeout 300 python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 2 --reps 5 --optimize $opt > gpurun_out/r3h/opt_$opt.log 2>&1
  echo "optimize $opt: $(grep 'rep 4' gpurun_out/r3h/opt_$opt.log | cut -c1-60) $(grep 'J sum' gpurun_out/r3h/opt_$opt.log)"
done
