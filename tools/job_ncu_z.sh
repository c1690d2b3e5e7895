TAG=$1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:z_kernel|col_kernel' -f -o /tmp/prof_$TAG python tools/prof_step.py --n 10 --batch 128 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full rc=$?"
python tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep gpurun_out/ncu_kernels_$TAG.csv > gpurun_out/ncu_kernels_$TAG.txt; echo "summary rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:z_kernel --launch-skip 1 --launch-count 1 2>/dev/null | gzip > gpurun_out/src_z10f_$TAG.csv.gz
ls -la gpurun_out | tail -5
