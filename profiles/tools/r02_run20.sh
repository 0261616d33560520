# 1 GPU: Blake3 leaf hash, TS_B3_ROT_FMA = 0..3 without the 64-register bound (3 CTAs per SM)
set -u
mkdir -p gpurun_out
for v in rot0 rot1 rot2 rot3 rot2; do
  export TAPSTARK_LIB=$PWD/tap-stark_b200/libtapstark_$v.so
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02t_$v.json 2> gpurun_out/r02t_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02t_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), {k:round(x["ms_per_step"],3) for k,x in d["stages"].items()}, d["self_check"]["root_match"], d["result"]["root"][:16])
except Exception as e:
    print("$v ERR", e); print(open("gpurun_out/r02t_$v.err").read()[-800:])
PY
done
