mkdir -p gpurun_out/r3d
export PYTHONPATH=$PWD
for ex in heat heat-explicit burgers bbm cahn-hilliard; do
  echo "=== examples/$ex.py" >> gpurun_out/r3d/examples.log
  ( time timeout 600 python examples/$ex.py ) >> gpurun_out/r3d/examples.log 2>&1; echo "exit code $?" >> gpurun_out/r3d/examples.log
done
grep -n "exit code\|real\|errornorm\|=== " gpurun_out/r3d/examples.log
