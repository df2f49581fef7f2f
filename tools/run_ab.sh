mkdir -p gpurun_out
python tools/ab_cycle.py gym_sbr2_b200/_variants/mb1.so gym_sbr2_b200/_variants/mb6.so gym_sbr2_b200/_variants/mb8.so > gpurun_out/ab_cycle_r01h.log 2>&1; cut -c1-400 gpurun_out/ab_cycle_r01h.log
