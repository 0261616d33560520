# 8 GPUs: end to end with the slab form of the host input at N=8 and N=4
set -u
mkdir -p gpurun_out
run() {
  env $3 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $2 --steps 5 --warmup 3 $4 > gpurun_out/r02q_$1.json 2> gpurun_out/r02q_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02q_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["ms_per_step"],3), d.get("e2e"), {k:round(v,2) for k,v in d["phases_last_step_ms"].items()}, d["self_check"]["root_match"], d["result"]["root"][:16])
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r02q_$1.err").read()[-1500:])
PY
}
run n8 8 TS_HOST_INPUT=slab "--no-self-check"
run n4 4 TS_HOST_INPUT=slab "--no-self-check"
