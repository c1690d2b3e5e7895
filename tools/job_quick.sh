python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01o.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r01o.log
python tools/conv_time.py > gpurun_out/conv_time.json 2> gpurun_out/conv_time.err; cat gpurun_out/conv_time.json
