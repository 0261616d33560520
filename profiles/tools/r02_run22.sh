# 1 GPU: the default bench line on the final build (as much as the remaining GPU budget allows)
mkdir -p gpurun_out
timeout 70 python bench.py --no-cpu-baseline > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02v_bench_n1.json").read().strip().splitlines()[-1])
print("bench", round(d["ms_per_step"],3), d["e2e"], d["roofline"]["frac"], d["gpu_launches"], d["self_check"], d["clocks"])
PY
