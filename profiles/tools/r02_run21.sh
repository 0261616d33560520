# 1 GPU: the full GPU suite on the final build (TS_B3_ROT_FMA=2 default)
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/r02u_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest_gpu.log; tail -3 gpurun_out/r02u_pytest_gpu.log
