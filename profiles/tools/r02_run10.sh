# 1 GPU: full GPU suite after the open / quotient / pageable / TapTreeMmcs changes, then A/B of the new switches
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02j_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest_gpu.log; tail -4 gpurun_out/r02j_pytest_gpu.log
python bench.py > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err; tail -c 400 gpurun_out/r02j_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02j_bench_n1.json").read().strip().splitlines()[-1])
print("bench", round(d["ms_per_step"],3), d["e2e"], d["stages"] if "stages" in d else None, d["self_check"])
PY
for t in 14 16 17 18 19; do
  echo "TS_TREE3_MIN_LOG=$t"
  TS_TREE3_MIN_LOG=$t python profiles/tools/config_sweep.py fri > gpurun_out/r02j_fri_tree3_$t.jsonl 2> gpurun_out/r02j_fri_tree3_$t.err; cut -c1-200 gpurun_out/r02j_fri_tree3_$t.jsonl
  TS_TREE3_MIN_LOG=$t python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-self-check > gpurun_out/r02j_bench_tree3_$t.json 2> gpurun_out/r02j_bench_tree3_$t.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02j_bench_tree3_$t.json").read().strip().splitlines()[-1])
print("  bench tree3_min_log=$t", round(d["ms_per_step"],3), {k: round(v["ms_per_step"],3) for k,v in d["stages"].items()})
PY
done
python profiles/tools/config_sweep.py open > gpurun_out/r02j_open_c4.jsonl 2> gpurun_out/r02j_open_c4.err; cut -c1-500 gpurun_out/r02j_open_c4.jsonl
for k in 0 2 4 8 16; do
  echo "TS_QV_CTAS_PER_SM=$k"
  TS_QV_CTAS_PER_SM=$k python profiles/tools/quotient_bench.py > gpurun_out/r02j_quotient_cps$k.jsonl 2> gpurun_out/r02j_quotient_cps$k.err; cut -c1-60,150-330 gpurun_out/r02j_quotient_cps$k.jsonl
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02j_open_launches_ncu.csv python profiles/tools/config_sweep.py open > gpurun_out/r02j_open_ncu.log 2>&1
tail -2 gpurun_out/r02j_open_ncu.log | cut -c1-300
