# 1 GPU: TMA staging (no swizzle) parity + A/B, occupancy variant, e2e incl. pageable source, ncu of the TMA kernel
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02e_build.log 2>&1
profiles/tools/tma_probe 1 > gpurun_out/r02e_tma_probe.jsonl 2>&1
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tma_staging" > gpurun_out/r02e_pytest_tma.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest_tma.log
tail -3 gpurun_out/r02e_pytest_tma.log
show() {
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02e_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["ms_per_step"],3), {k:round(x["ms_per_step"],2) for k,x in d["stages"].items()}, d["self_check"] and d["self_check"]["root_match"], d["result"]["root"][:16], d.get("e2e"))
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r02e_$1.err").read()[-1500:])
PY
}
if grep -q "rc=0" gpurun_out/r02e_pytest_tma.log; then
  TS_TMA=1 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02e_tma.json 2> gpurun_out/r02e_tma.err; show tma
  TS_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_tma -c 2 -o gpurun_out/r02e_tma python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-self-check > gpurun_out/r02e_ncu.log 2>&1
fi
TAPSTARK_LIB=$PWD/tap-stark_b200/libtapstark_mb2.so python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02e_mb2.json 2> gpurun_out/r02e_mb2.err; show mb2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_full.json 2> gpurun_out/r02e_full.err; show full
