# GPU suite of the final build (incl. the bit-identity test of the kernel switches) and ncu traffic captures for the
# dominant kernels of the two sub-records (S_9 x 1024 fp32: top-level T kernel; S_10 fp64: level-8 column kernel)
TAG=$1
COMMIT=$2
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:t_kernel' -f -o /tmp/prof_s9 python tools/prof_step.py --n 9 --batch 1024 > gpurun_out/ncu_s9_$TAG.log 2>&1; echo "ncu s9 rc=$?"
python tools/ncu_summary.py /tmp/prof_s9.ncu-rep gpurun_out/ncu_kernels_${TAG}_s9.csv > /dev/null
SNFFT_COMMIT=$COMMIT python tools/ncu_traffic.py gpurun_out/ncu_kernels_${TAG}_s9.csv 1024 f32 gpurun_out/${TAG}_traffic_s9.json 9 > /dev/null
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:col_kernel' -f -o /tmp/prof_f64 python tools/prof_step.py --n 10 --batch 128 --dtype f64 > gpurun_out/ncu_f64_$TAG.log 2>&1; echo "ncu f64 rc=$?"
python tools/ncu_summary.py /tmp/prof_f64.ncu-rep gpurun_out/ncu_kernels_${TAG}_f64.csv > /dev/null
SNFFT_COMMIT=$COMMIT python tools/ncu_traffic.py gpurun_out/ncu_kernels_${TAG}_f64.csv 128 f64 gpurun_out/${TAG}_traffic_f64.json 10 > /dev/null
cut -d, -f1-6 gpurun_out/ncu_kernels_${TAG}_s9.csv gpurun_out/ncu_kernels_${TAG}_f64.csv
cat gpurun_out/${TAG}_traffic_s9.json gpurun_out/${TAG}_traffic_f64.json | head -60
