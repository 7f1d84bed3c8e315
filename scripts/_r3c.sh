python -m pytest tests -m gpu -x -q > gpurun_out/r3c_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r3c_tests.log
python bench.py > gpurun_out/r3c_bench_n1.json 2> gpurun_out/r3c_bench_n1.err; echo bench rc=$?; tail -c 600 gpurun_out/r3c_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3c_bench_ref.json 2> gpurun_out/r3c_bench_ref.err; echo ref rc=$?; tail -c 300 gpurun_out/r3c_bench_ref.err; head -c 600 gpurun_out/r3c_bench_ref.json
