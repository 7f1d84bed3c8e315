for ex in p2p all_to_all; do
ENRGY_SHADE_EXCHANGE=$ex python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --only c3_shadow > gpurun_out/r3a_n8_$ex.json 2> gpurun_out/r3a_n8_$ex.err; echo rc=$?
python - <<EOF
import json
d=json.load(open('gpurun_out/r3a_n8_$ex.json'))
c=d['configs']['c3_shadow']; print('$ex', c['ms_per_step'], c['value'], c['sweep_ms_per_rank'], c['fused_ms_per_rank'], c['exchange'], c['chunks_per_pass'])
EOF
done
