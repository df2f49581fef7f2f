mkdir -p gpurun_out
NG=${NG:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_n${NG}_r01h.log 2> gpurun_out/bench_n${NG}_r01h.err; echo rc=$?; tail -c 300 gpurun_out/bench_n${NG}_r01h.err; python - <<PY
import json
for l in open('gpurun_out/bench_n${NG}_r01h.log'):
    if l.startswith('{"metric"'):
        d = json.loads(l); print(d['value'], d['n_gpus'], d['roofline']['frac'], d['e2e']['value']); print(json.dumps(d['paths'].get('config5_rollout'))[:900])
PY
