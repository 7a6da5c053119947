ncu --set full --clock-control none -k regex:'msm_accumulate' -s 1 -c 1 -o /tmp/r2F_fat -f python tools/ncu_pair_target.py > gpurun_out/r2F_ncu.log 2>&1
ncu -i /tmp/r2F_fat.ncu-rep --page raw --csv > gpurun_out/r2F_fat_raw.csv 2>/dev/null
tail -2 gpurun_out/r2F_ncu.log
