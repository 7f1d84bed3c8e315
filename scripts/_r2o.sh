python -m pytest tests -m gpu -x -q -k "stations" > gpurun_out/r2o_tests.log 2>&1; echo tests rc=$?; tail -5 gpurun_out/r2o_tests.log
python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 > gpurun_out/r2o_plain.log 2>&1
ENRGY_B200_LIB=scratch/variants/lw_estrin.so python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 >> gpurun_out/r2o_plain.log 2>&1
python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 >> gpurun_out/r2o_plain.log 2>&1
ENRGY_B200_LIB=scratch/variants/lw_estrin.so python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 >> gpurun_out/r2o_plain.log 2>&1
cat gpurun_out/r2o_plain.log
