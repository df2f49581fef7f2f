mkdir -p gpurun_out
python tools/os_curve.py dp45 > gpurun_out/os_curve_r02a.log 2>&1; tail -2 gpurun_out/os_curve_r02a.log
