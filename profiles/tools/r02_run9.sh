# 1 GPU, final: smoke, full GPU test suite, bench (default flags), reference arm, ncu launch list, sweeps
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02i_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02i_smoke.log; tail -2 gpurun_out/r02i_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest_gpu.log; tail -4 gpurun_out/r02i_pytest_gpu.log
python bench.py > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; tail -c 600 gpurun_out/r02i_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02i_bench_n1.json").read().strip().splitlines()[-1])
print("bench", round(d["ms_per_step"],3), d["e2e"], d["roofline"]["frac"], d["lde_stage"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["self_check"])
PY
python bench.py --impl reference --steps 3 --warmup 1 --cpu-same-config > gpurun_out/r02i_bench_ref.json 2> gpurun_out/r02i_bench_ref.err; cut -c1-700 gpurun_out/r02i_bench_ref.json
python profiles/tools/config_sweep.py > gpurun_out/r02i_config_sweep.jsonl 2> gpurun_out/r02i_config_sweep.err; cat gpurun_out/r02i_config_sweep.jsonl | cut -c1-400
python profiles/tools/quotient_bench.py > gpurun_out/r02i_quotient.jsonl 2> gpurun_out/r02i_quotient.err; cat gpurun_out/r02i_quotient.jsonl | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02i_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-self-check > gpurun_out/r02i_ncu_launches.log 2>&1
tail -2 gpurun_out/r02i_ncu_launches.log | cut -c1-300
