mkdir -p gpurun_out
for v in d4 d6; do
SBR_B200_LIB=$PWD/gym_sbr2_b200/_variants/$v.so python bench.py --mode dp45 --rtol 1e-6 --atol 1e-8 --steps 3 --warmup 3 --no-cpu-baseline --no-interval-path --no-rollout 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric\"'):
        d = json.loads(l); print('$v', d['value'], d['roofline']['kernel_ms'], d['roofline'].get('rhs_per_env'))
"
done
