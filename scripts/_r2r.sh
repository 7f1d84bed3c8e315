python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --only c3_shadow > gpurun_out/r2r_n8.json 2> gpurun_out/r2r_n8.err; echo rc=$?
python - <<EOF
import json
d=json.load(open('gpurun_out/r2r_n8.json'))
c=d['configs']['c3_shadow']; print(c['ms_per_step'], c['value'], c['sweep_ms_per_rank'], c['fused_ms_per_rank'], c['exchange'], c['chunks_per_pass'], c['band_rows'])
print(d['value'], d['e2e']['value'])
EOF
