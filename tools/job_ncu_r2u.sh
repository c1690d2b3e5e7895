TAG=$1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:t_kernel|col_kernel' -f -o /tmp/prof_$TAG python tools/prof_step.py --n 10 --batch 128 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full rc=$?"
python tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep gpurun_out/ncu_kernels_$TAG.csv > gpurun_out/ncu_kernels_$TAG.txt; echo "summary rc=$?"
python tools/ncu_hot.py /tmp/prof_$TAG.ncu-rep 40 > gpurun_out/ncu_hot_$TAG.txt 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page details --csv --kernel-name-base demangled 2>/dev/null | grep -i 'L2\|DRAM\|sector\|Coalesc\|uncoalesced\|excessive' | cut -c1-400 | head -150 > gpurun_out/ncu_details_$TAG.txt
ls -la gpurun_out | tail -5
