python bench.py --no-cpu-baseline > gpurun_out/r3h_bench_n1.json 2> gpurun_out/r3h_bench_n1.err; echo bench rc=$?; tail -c 300 gpurun_out/r3h_bench_n1.err
