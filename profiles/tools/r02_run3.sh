# 8 GPUs: scaling bench (copy-engine re-shard), N=8 NCCL re-shard for comparison
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02c_build.log 2>&1
run() {  # name, nproc, extra env, extra args
  env $3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $2 --steps 6 --warmup 3 $4 > gpurun_out/r02c_$1.json 2> gpurun_out/r02c_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["ms_per_step"],3), (d.get("e2e") or {}).get("ms_per_step"), d["phases_last_step_ms"], d["self_check"], d["result"]["root"][:16], {k:round(v["ms_per_step"],2) for k,v in d["stages"].items()})
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r02c_$1.err").read()[-1500:])
PY
}
run n8_ce 8 TS_RESHARD=ce ""
run n8_nccl 8 TS_RESHARD=nccl "--no-e2e"
run n4_ce 4 TS_RESHARD=ce ""
