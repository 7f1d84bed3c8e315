# launch list of the default bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; echo launches rc=$?
NCU="ncu --set full --clock-control none --import-source on -c 1"
run() { name=$1; shift; skip=$1; shift; kern=$1; shift
  $NCU -k regex:$kern -s $skip -o /tmp/r02_$name -f "$@" > gpurun_out/r02_ncu_$name.log 2>&1; echo $name rc=$?
  python scripts/summarize_ncu.py /tmp/r02_$name.ncu-rep gpurun_out/r02_ncu_$name.csv; }
run f32 2 energy_balance_kernel python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200
run f64 2 energy_balance_kernel python scripts/measure_shadow.py --noshadow --size 2048 --nsteps 2200 --dtype f64
run sweep 2 sweep_kernel python scripts/measure_shadow.py --size 4096 --nsteps 96
run masked 1 energy_balance_kernel python scripts/measure_shadow.py --size 4096 --nsteps 96
run msm 2 energy_balance_kernel python scripts/measure_modes.py msm --nsteps 512
run members 2 energy_balance_kernel python scripts/measure_modes.py ensemble --members 4 --size 2048 --nsteps 512 --nostats
ls -la gpurun_out/
