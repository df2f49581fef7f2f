mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01f.log 2>&1; tail -1 gpurun_out/smoke_r01f.log
python bench.py > gpurun_out/bench_r01f.log 2> gpurun_out/bench_r01f.err; echo bench rc=$?
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-interval-path --no-rollout > gpurun_out/bench_plain_r01f.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-interval-path --no-rollout > gpurun_out/ncu_list_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sbr_cycle_v2 -s 2 -c 1 -f -o gpurun_out/prof_cycle_rk4_r01f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-interval-path --no-rollout > gpurun_out/ncu_full_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sbr_cycle_v2 -s 2 -c 1 -f -o gpurun_out/prof_cycle_dp45_ordered_r01f python bench.py --mode dp45 --rtol 1e-6 --atol 1e-8 --steps 1 --warmup 3 --no-cpu-baseline --no-interval-path --no-rollout >> gpurun_out/ncu_full_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sbr_v4_step -s 120 -c 1 -f -o gpurun_out/prof_v4_step_dp45_r01f python tools/prof_v4.py 130 >> gpurun_out/ncu_full_f.log 2>&1
ls -la gpurun_out/*r01f*; tail -c 300 gpurun_out/bench_r01f.err
