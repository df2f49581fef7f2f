mkdir -p gpurun_out
python tools/bench_os_variants.py gym_sbr2_b200/_variants/pf0.so gym_sbr2_b200/_variants/pf1.so gym_sbr2_b200/_variants/pf2.so > gpurun_out/os_ab_r01g.log 2>&1; cat gpurun_out/os_ab_r01g.log
python tools/v4_ab.py gym_sbr2_b200/_variants/pf0.so gym_sbr2_b200/_variants/pf1.so gym_sbr2_b200/_variants/pf2.so > gpurun_out/v4_ab_r01g.log 2>&1; cat gpurun_out/v4_ab_r01g.log
