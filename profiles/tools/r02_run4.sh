# 1 GPU: TMA staging and shuffle exchange -- parity, A/B bench, ncu of the TMA kernel
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02d_build.log 2>&1
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "tma_staging or shuffle_exchange or round_forms or pcs_open or commit_phase" > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -5 gpurun_out/r02d_pytest.log
for v in base tma shfl; do
  unset TS_TMA TS_SHFL
  [ $v = tma ] && export TS_TMA=1
  [ $v = shfl ] && export TS_SHFL=1
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02d_$v.json 2> gpurun_out/r02d_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02d_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), {k:round(x["ms_per_step"],2) for k,x in d["stages"].items()}, d["self_check"]["root_match"], d["result"]["root"][:16])
except Exception as e:
    print("$v ERR", e); print(open("gpurun_out/r02d_$v.err").read()[-1500:])
PY
done
unset TS_SHFL
export TS_TMA=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_tma -c 2 -o gpurun_out/r02d_tma python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-self-check > gpurun_out/r02d_ncu.log 2>&1
