# 2 GPUs: multi-GPU parity after the fused fold + leaf hash in the sharded FRI rounds and the new tree switch point; N=2 bench
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02l_build.log 2>&1
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -v > gpurun_out/r02l_pytest_sharded_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02l_pytest_sharded_2gpu.log
tail -6 gpurun_out/r02l_pytest_sharded_2gpu.log
run() {
  env $3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $2 --steps 5 --warmup 3 $4 > gpurun_out/r02l_$1.json 2> gpurun_out/r02l_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02l_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["ms_per_step"],3), (d.get("e2e") or {}).get("ms_per_step"), {k:round(v,2) for k,v in d["phases_last_step_ms"].items()}, d["self_check"]["root_match"], d["result"]["root"][:16], {k:round(v["ms_per_step"],2) for k,v in d["stages"].items()})
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r02l_$1.err").read()[-1500:])
PY
}
run n2 2 TS_X=0 ""
run n2_unfused 2 TS_NO_FOLD_HASH=1 "--no-e2e"
timeout 300 python profiles/tools/config_sweep.py open > gpurun_out/r02l_open_c4.jsonl 2> gpurun_out/r02l_open_c4.err; cut -c1-400 gpurun_out/r02l_open_c4.jsonl
