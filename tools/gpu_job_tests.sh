mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r2_final_pytest_gpu.log
