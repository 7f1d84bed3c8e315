python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/r2u_tests.log
python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 >> gpurun_out/r2u_plain.log 2>&1
python scripts/measure_shadow.py --size 4096 --nsteps 384 >> gpurun_out/r2u_plain.log 2>&1
cat gpurun_out/r2u_plain.log
