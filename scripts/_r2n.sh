python -m pytest tests -m gpu -x -q -k "stations" > gpurun_out/r2n_tests.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2n_tests.log
python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 > gpurun_out/r2n_plain.log 2>&1; cat gpurun_out/r2n_plain.log
for opt in "--sequential" "" "--nostats"; do python scripts/measure_modes.py ensemble --members 8 --size 4096 --nsteps 2200 $opt >> gpurun_out/r2n_ens.log 2>&1; done; cat gpurun_out/r2n_ens.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests_all.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2n_tests_all.log
