for ex in all_to_all copy; do
ENRGY_SHADE_EXCHANGE=$ex python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --only c3_shadow --c3-t 384 > gpurun_out/r3f_n2_$ex.json 2> gpurun_out/r3f_n2_$ex.err; echo rc=$?; tail -c 300 gpurun_out/r3f_n2_$ex.err
python - <<EOF
import json
d=json.load(open('gpurun_out/r3f_n2_$ex.json'))
c=d['configs']['c3_shadow']; print('$ex', c.get('ms_per_step'), c.get('value'), c.get('exchange'), c.get('chunks_per_pass'), c.get('check'), c.get('error'), c.get('roofline_sweep',{}).get('issue_slots',{}).get('frac'))
EOF
done
