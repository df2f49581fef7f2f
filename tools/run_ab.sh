mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -2 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo rc=$?; python - <<'PY'
import json
for l in open('gpurun_out/bench_final.log'):
    if l.startswith('{"metric"'):
        d = json.loads(l); p = d['paths']
        print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])
        print(json.dumps(p['sbr_v2_dp45']['rtol1e-06_env_order'])[:700])
PY
