TAG=$1
COMMIT=$2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:z_kernel|t_kernel|col_kernel|s5_kernel|gather_vec|pack_vec' -f -o /tmp/prof_$TAG python tools/prof_step.py --n 10 --batch 128 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full rc=$?"
python tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep gpurun_out/ncu_kernels_$TAG.csv > gpurun_out/ncu_kernels_$TAG.txt; echo "summary rc=$?"
SNFFT_COMMIT=$COMMIT python tools/ncu_traffic.py gpurun_out/ncu_kernels_$TAG.csv 128 f32 gpurun_out/${TAG}_traffic.json > /dev/null
cut -d, -f1-6 gpurun_out/ncu_kernels_$TAG.csv
# SASS-level source pages of the two kernel types that dominate (inverse level 8 column kernel, forward level 10 T kernel)
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:t_kernel --launch-skip 1 --launch-count 1 2>/dev/null | gzip > gpurun_out/src_t10_$TAG.csv.gz
ls -la gpurun_out/ | tail -12
