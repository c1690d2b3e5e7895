set -x
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fwd_tb_kernel<float, .int.512, .int.1, .int.7, .int.6' --launch-skip 3 -c 1 -f -o /tmp/prof_tb_k10 python tools/prof_step.py --n 10 --batch 16 > gpurun_out/ncu_tb.log 2>&1; echo "full rc=$?"
ncu -i /tmp/prof_tb_k10.ncu-rep --page source --csv --kernel-name-base demangled | cut -d, -f1-10 | gzip > gpurun_out/tb_k10_src.csv.gz
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:inv_tb_kernel<float, .int.512, .int.1, .int.7, .int.6' --launch-skip 3 -c 1 -f -o /tmp/prof_tb_k10i python tools/prof_step.py --n 10 --batch 16 > gpurun_out/ncu_tb.log 2>&1; echo "full rc=$?"
ncu -i /tmp/prof_tb_k10i.ncu-rep --page source --csv --kernel-name-base demangled | cut -d, -f1-10 | gzip > gpurun_out/tb_k10i_src.csv.gz
ls -la gpurun_out
