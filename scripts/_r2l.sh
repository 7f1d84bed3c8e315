python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests_all.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2l_tests_all.log
python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 > gpurun_out/r2l_plain.log 2>&1; cat gpurun_out/r2l_plain.log
