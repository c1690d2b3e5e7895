timeout 170 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:col_kernel|s5_kernel' -f -o /tmp/prof_col python tools/prof_step.py --n 10 --batch 128 > gpurun_out/ncu_col_r01r.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/prof_col.ncu-rep gpurun_out/ncu_col_r01r.csv > /dev/null; echo "summary rc=$?"
