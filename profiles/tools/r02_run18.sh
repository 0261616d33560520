# 1 GPU: Pcs::open at C4 after routing narrow matrices to the scalar barycentric kernel; smoke
set -u
mkdir -p gpurun_out
python profiles/tools/config_sweep.py open > gpurun_out/r02r_open_c4.jsonl 2> gpurun_out/r02r_open_c4.err; cut -c1-400 gpurun_out/r02r_open_c4.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
