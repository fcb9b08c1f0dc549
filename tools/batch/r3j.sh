mkdir -p gpurun_out/r3j
timeout 600 python tools/time_config.py -n 184 --tile 0 --scatter 2 --reps 4 > gpurun_out/r3j/lex.log 2>&1
echo "lexicographic: $(grep 'rep 3' gpurun_out/r3j/lex.log | cut -c1-60); $(grep '^plan' gpurun_out/r3j/lex.log | cut -c1-140); $(tail -2 gpurun_out/r3j/lex.log | cut -c1-200)"
timeout 600 python tools/time_config.py -n 184 --tile 0 --scatter 1 --reps 4 > gpurun_out/r3j/lex_atomic.log 2>&1
echo "lexicographic atomic: $(grep 'rep 3' gpurun_out/r3j/lex_atomic.log | cut -c1-60)"
