timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 --profile-host > gpurun_out/r2_b2_c19.json 2> gpurun_out/r2_b2_c19.err
grep '^{' gpurun_out/r2_b2_c19.json | python -c "
import sys, json
d=json.loads(sys.stdin.readline())
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['host_enqueue_ms_per_step'], d.get('host_enqueue_ms_per_step'), d.get('comm'))
"
grep -A45 "cumulative" gpurun_out/r2_b2_c19.err | head -80
