python -m pytest tests -m gpu -x -q -k "ensemble" > gpurun_out/r2k_tests.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2k_tests.log
for opt in "" "--nostats" "--shadow" "--shadow --nostats" "--dtype f64 --nostats"; do python scripts/measure_modes.py ensemble --members 8 --size 4096 --nsteps 384 $opt >> gpurun_out/r2k_ens.log 2>&1; done
for opt in "" "--nostats" "--shadow --nostats"; do ENRGY_B200_LIB=scratch/variants/mem_minb4.so python scripts/measure_modes.py ensemble --members 8 --size 4096 --nsteps 384 $opt >> gpurun_out/r2k_ens.log 2>&1; done
cat gpurun_out/r2k_ens.log
python scripts/measure_shadow.py --size 4096 --nsteps 384 >> gpurun_out/r2k_shadow.log 2>&1; cat gpurun_out/r2k_shadow.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2k_tests_all.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2k_tests_all.log
