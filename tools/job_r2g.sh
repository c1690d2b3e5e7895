TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_$TAG.log
python tools/sanitize_cases.py > gpurun_out/sanitize_plain_$TAG.log 2>&1; echo "cases rc=$?"; tail -3 gpurun_out/sanitize_plain_$TAG.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_cases.py > gpurun_out/memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|SANITIZE_CASES_OK" gpurun_out/memcheck_$TAG.log
