for cfg in "0 3 0" "1 3 0" "1 4 0" "0 4 32"; do set -- $cfg; 
 if [ "$1" = "1" ]; then export ZK_BA_NOTREE=1; else unset ZK_BA_NOTREE; fi
 export ZK_BA_MINB=$2 ZK_BA_K=$3
 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_l_$1_$2_$3.csv python tools/kernel_times.py msm20 > /dev/null 2>&1
done
