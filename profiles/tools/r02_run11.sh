# 1 GPU: GPU suite after fold_hash / inv_denoms / bary / dot_small, A/B of the fused fold + leaf hash, open at C4
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02k_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest_gpu.log; tail -4 gpurun_out/r02k_pytest_gpu.log
for v in fused unfused; do
  if [ $v = unfused ]; then export TS_NO_FOLD_HASH=1; else unset TS_NO_FOLD_HASH; fi
  python profiles/tools/config_sweep.py fri > gpurun_out/r02k_fri_$v.jsonl 2> gpurun_out/r02k_fri_$v.err; echo "fri $v"; cut -c1-120 gpurun_out/r02k_fri_$v.jsonl
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-self-check > gpurun_out/r02k_bench_$v.json 2> gpurun_out/r02k_bench_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02k_bench_$v.json").read().strip().splitlines()[-1])
print("  bench $v", round(d["ms_per_step"],3), {k: (round(v["ms_per_step"],3), v["launches_per_step"]) for k,v in d["stages"].items()})
PY
done
unset TS_NO_FOLD_HASH
python profiles/tools/config_sweep.py open > gpurun_out/r02k_open_c4.jsonl 2> gpurun_out/r02k_open_c4.err; cut -c1-400 gpurun_out/r02k_open_c4.jsonl
python bench.py > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; tail -c 300 gpurun_out/r02k_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02k_bench_n1.json").read().strip().splitlines()[-1])
print("bench", round(d["ms_per_step"],3), d["e2e"], d["roofline"], d["gpu_launches"], d["self_check"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02k_open_launches_ncu.csv python profiles/tools/config_sweep.py open > gpurun_out/r02k_open_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02k_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-self-check > gpurun_out/r02k_ncu_launches.log 2>&1
tail -1 gpurun_out/r02k_ncu_launches.log | cut -c1-200
