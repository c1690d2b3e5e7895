ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gather_kernel|pack_kernel|l6_kernel|s5_kernel' -c 8 -f -o /tmp/prof_dm python tools/prof_step.py --n 10 --batch 128 > gpurun_out/ncu_dm.log 2>&1; echo "full rc=$?"
python tools/ncu_summary.py /tmp/prof_dm.ncu-rep gpurun_out/dm.csv > /dev/null
python tools/ncu_hot.py /tmp/prof_dm.ncu-rep 25 > gpurun_out/dm_hot.txt
