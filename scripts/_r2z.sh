python -m pytest tests -m gpu -x -q -k "shading or c3_full or smoke" > gpurun_out/r2z_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/r2z_tests.log
for lib in "" scratch/variants/old_sweep.so; do ENRGY_B200_LIB=$lib python scripts/measure_shadow.py --size 4096 --nsteps 384 >> gpurun_out/r2z_plain.log 2>&1; ENRGY_B200_LIB=$lib python scripts/measure_shadow.py --size 8192 --nsteps 96 >> gpurun_out/r2z_plain.log 2>&1; done
cat gpurun_out/r2z_plain.log
