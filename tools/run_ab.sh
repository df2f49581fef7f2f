mkdir -p gpurun_out
python tools/v4_ab.py gym_sbr2_b200/_variants/v4_mb6.so gym_sbr2_b200/_variants/v4_mb5.so gym_sbr2_b200/_variants/v4_mb4.so gym_sbr2_b200/_variants/v4_mb8.so > gpurun_out/v4_ab_r02a.log 2>&1; cat gpurun_out/v4_ab_r02a.log
