# strong-scaling spot check on one 8-GPU box: the fixed sweep at N = 8, 4, 2 (no secondary / weak / CPU legs)
for N in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-weak --no-cpu-baseline > gpurun_out/r2_strong_n${N}.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_strong_n${N}.json') if l.startswith('{')][-1]); print('N=$N', round(d['value']/1e6,3), 'M', round(d['ms_per_step'],1), 'ms', d['clocks']['sm_mhz'], d['clocks']['power_w'])"
done
