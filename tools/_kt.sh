ncu --set full --clock-control none -k regex:"k_rowcol_sums" -s 8 -c 1 -o /tmp/r2_rc -f python tools/prove_bench.py gpu 256 > gpurun_out/r2_ncu3.log 2>&1
python tools/ncu_summary.py /tmp/r2_rc.ncu-rep gpurun_out/r02_ncu_full
ncu --set full --clock-control none -k regex:"k_ba_backward" -s 8 -c 1 -o /tmp/r2_bb -f python tools/prove_bench.py gpu 256 > gpurun_out/r2_ncu4.log 2>&1
python tools/ncu_summary.py /tmp/r2_bb.ncu-rep gpurun_out/r02_ncu_full
