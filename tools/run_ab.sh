mkdir -p gpurun_out
python tools/ab_cycle.py gym_sbr2_b200/_variants/a_base.so gym_sbr2_b200/_variants/b_tab.so > gpurun_out/ab_cycle_r02b.log 2>&1; cat gpurun_out/ab_cycle_r02b.log
python tools/bench_os_variants.py gym_sbr2_b200/_variants/a_base.so gym_sbr2_b200/_variants/b_tab.so > gpurun_out/os_ab_r02b.log 2>&1; cat gpurun_out/os_ab_r02b.log
