# 1 GPU: Blake3 leaf hash with some rotations on the fma pipe (TS_B3_ROT_FMA = 2 / 4 of 8 G per round), with and without the 64-register bound
set -u
mkdir -p gpurun_out
for v in base rot2 rot4 rot2mb4 rot4mb4; do
  if [ $v = base ]; then unset TAPSTARK_LIB; else export TAPSTARK_LIB=$PWD/tap-stark_b200/libtapstark_$v.so; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02s_$v.json 2> gpurun_out/r02s_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02s_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), {k:round(x["ms_per_step"],3) for k,x in d["stages"].items()}, d["self_check"]["root_match"], d["result"]["root"][:16])
except Exception as e:
    print("$v ERR", e); print(open("gpurun_out/r02s_$v.err").read()[-800:])
PY
done
