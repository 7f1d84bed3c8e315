python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo rc=$?; tail -c 1200 gpurun_out/r2p_bench.err
python -m pytest tests -m gpu -x -q -k "dropin" > gpurun_out/r2p_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2p_tests.log
