# 1 GPU, final: smoke, full GPU suite, default bench, open at C4 (+ launch list)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02o_smoke.log; tail -2 gpurun_out/r02o_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02o_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest_gpu.log; tail -4 gpurun_out/r02o_pytest_gpu.log
python bench.py > gpurun_out/r02o_bench_n1.json 2> gpurun_out/r02o_bench_n1.err; tail -c 300 gpurun_out/r02o_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02o_bench_n1.json").read().strip().splitlines()[-1])
print("bench", round(d["ms_per_step"],3), d["e2e"], d["roofline"]["frac"], d["gpu_launches"], d["self_check"], d["clocks"])
PY
python profiles/tools/config_sweep.py open > gpurun_out/r02o_open_c4.jsonl 2> gpurun_out/r02o_open_c4.err; cut -c1-400 gpurun_out/r02o_open_c4.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02o_open_launches_ncu.csv python profiles/tools/config_sweep.py open > gpurun_out/r02o_open_ncu.log 2>&1
python profiles/tools/quotient_bench.py > gpurun_out/r02o_quotient.jsonl 2> gpurun_out/r02o_quotient.err; cut -c1-60,150-330 gpurun_out/r02o_quotient.jsonl
