ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_msm20_fresh.csv python tools/kernel_times.py msm20 fresh > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_msm20.csv python tools/kernel_times.py msm20 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_prove256.csv python tools/prove_bench.py gpu 256 > gpurun_out/r2_prove256.log 2>&1
tail -3 gpurun_out/r2_prove256.log
