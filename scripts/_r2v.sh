python -m pytest tests -m gpu -x -q -k "shading or full_size or smoke or ensemble or stations" > gpurun_out/r2v_tests.log 2>&1; echo tests rc=$?; tail -4 gpurun_out/r2v_tests.log
for lib in "" scratch/variants/ring0.so ""; do ENRGY_B200_LIB=$lib python scripts/measure_shadow.py --size 4096 --nsteps 384 >> gpurun_out/r2v_plain.log 2>&1; done
ENRGY_B200_LIB= python scripts/measure_shadow.py --size 2048 --nsteps 256 >> gpurun_out/r2v_plain.log 2>&1
cat gpurun_out/r2v_plain.log
