mkdir -p gpurun_out/r2f
for tag in "" _m8 _m10 _t256m4 _t64m16; do
  export FDTS_B200_LIB=$PWD/firedrake_ts_b200/csrc/libfdts_b200_dev$tag.so
  for dbg in 0 128 144; do
    python tools/time_config.py -n 184 --tile 3 --ftile 6 --scatter 1 --reps 4 --debug $dbg > gpurun_out/r2f/resid${tag}_dbg_$dbg.log 2>&1
    echo "variant '$tag' dbg $dbg: $(grep 'rep 3' gpurun_out/r2f/resid${tag}_dbg_$dbg.log | awk '{print $4}')"
  done
done
