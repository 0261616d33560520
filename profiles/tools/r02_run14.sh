# 8 GPUs: final N=8 / N=4 lines (fused fold + leaf hash in the sharded rounds, tree switch point 2^19) and 2 chunks per rank at N=8
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02n_build.log 2>&1
run() {
  env $3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $2 --steps 5 --warmup 3 $4 > gpurun_out/r02n_$1.json 2> gpurun_out/r02n_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02n_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["ms_per_step"],3), (d.get("e2e") or {}).get("ms_per_step"), {k:round(v,2) for k,v in d["phases_last_step_ms"].items()}, d["self_check"]["root_match"], d["result"]["root"][:16], {k:round(v["ms_per_step"],2) for k,v in d["stages"].items()})
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r02n_$1.err").read()[-1500:])
PY
}
run n8 8 TS_X=0 ""
run n8_chunks2 8 TS_SHARD_CHUNKS=2 "--no-e2e"
run n4 4 TS_X=0 ""
