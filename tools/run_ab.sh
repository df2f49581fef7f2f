mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01h.log 2> gpurun_out/bench_r01h.err; echo bench rc=$?; tail -c 300 gpurun_out/bench_r01h.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01h.log 2>&1; tail -1 gpurun_out/bench_ref_r01h.log | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/bench_r01h.log'):
    if l.startswith('{"metric"'):
        d = json.loads(l); p = d['paths']
        print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['cpu_baseline']['value'])
        print(p['sbros_v1']['dp45']['ms_per_episode'], p['sbros_v1']['dp45']['interval_steps_per_sec'], p['sbros_v1']['dp45']['ms_per_plain_step'], p['sbros_v1']['cpu_baseline']['value'])
        print(p['sbr_v2_dp45']['rtol1e-06_ordered']['ms'], p['sbr_v2_rk4_7substeps']['kernel_ms'], p['sbr_v4']['ms_per_episode'], p['config5_rollout']['ms_episode'], p['config1_small_batch'])
PY
