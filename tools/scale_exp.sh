for cfg in "8 8" "8 4" "4 8" "4 4"; do set -- $cfg; N=$1; M=$2;
PP_BENCH_MICRO=$M timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$M bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-weak --no-cpu-baseline > gpurun_out/r2_exp_n${N}_m${M}.json 2>/dev/null
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_exp_n${N}_m${M}.json') if l.startswith('{')][-1]); print('N=$N MICRO=$M', round(d['value']/1e6,3), 'M', round(d['ms_per_step'],1), 'ms', d['clocks']['sm_mhz'], d['clocks']['power_w'])"
done
