set -u
bash profiles/tools/gpu_ab.sh r02a
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_kernel -c 8 -o gpurun_out/r02a_v4 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_pytest.log
python - <<'PY'
import json
for v in ("v4","nov4"):
    try:
        d=json.loads(open(f"gpurun_out/r02a_{v}.json").read().strip().splitlines()[-1])
        print(v, d["ms_per_step"], {k:round(x["ms_per_step"],2) for k,x in d["stages"].items()})
    except Exception as e: print(v, "ERR", e)
PY
