mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01g.log 2>&1; tail -2 gpurun_out/pytest_gpu_r01g.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-interval-path --no-rollout > gpurun_out/bench_plain_r01g.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/bench_plain_r01g.log'):
    if l.startswith('{"metric"'):
        d = json.loads(l); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'])
PY
