# 1 GPU: ncu --set full of the barycentric kernels inside Pcs::open at the C4 shape
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02m_build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bary_partial -c 3 -o gpurun_out/r02m_bary -f python profiles/tools/config_sweep.py open > gpurun_out/r02m_ncu.log 2>&1
tail -3 gpurun_out/r02m_ncu.log | cut -c1-200
ls -la gpurun_out/r02m_bary.ncu-rep
TS_NO_BARY4=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:bary_partial -c 1 -o gpurun_out/r02m_bary_scalar -f python profiles/tools/config_sweep.py open > gpurun_out/r02m_ncu2.log 2>&1
ls -la gpurun_out/r02m_bary_scalar.ncu-rep
