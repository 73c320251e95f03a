# 2 GPUs: stop-head join deferred past the flag exchange and GAE (1) against joined after the rollout (0), same box
for v in 0 1 0 1; do
PLUME_EXPERIMENT_DEFER_JOIN=$v timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29560+RANDOM%200)) bench.py --gpus 2 --steps 20 --warmup 3 --skip-aux --skip-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('defer=$v', round(d['ms_per_step'],3), 'ms', d['value'])"
done
