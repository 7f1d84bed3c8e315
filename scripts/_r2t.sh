python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/r2t_tests.log
for lib in "" scratch/variants/onebar0.so; do ENRGY_B200_LIB=$lib python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 >> gpurun_out/r2t_plain.log 2>&1; done
for lib in "" scratch/variants/rcp0.so; do ENRGY_B200_LIB=$lib python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 --dtype f64 >> gpurun_out/r2t_plain.log 2>&1; done
for lib in "" scratch/variants/ma8.so scratch/variants/ma16.so scratch/variants/onebar0.so; do ENRGY_B200_LIB=$lib python scripts/measure_shadow.py --size 4096 --nsteps 384 >> gpurun_out/r2t_plain.log 2>&1; done
cat gpurun_out/r2t_plain.log
