# 2 GPUs: multi-GPU parity tests (committed log) + N=2 bench with the three re-shard forms
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02b_build.log 2>&1
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -v > gpurun_out/r02b_pytest_sharded_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest_sharded_2gpu.log
tail -8 gpurun_out/r02b_pytest_sharded_2gpu.log
for mode in ce nccl; do
  TS_RESHARD=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02b_n2_$mode.json 2> gpurun_out/r02b_n2_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02b_n2_$mode.json").read().strip().splitlines()[-1])
    print("$mode", d["ms_per_step"], d.get("e2e",{}).get("ms_per_step"), d["phases_last_step_ms"], d["self_check"], d["result"]["root"][:16])
except Exception as e:
    print("$mode ERR", e); print(open("gpurun_out/r02b_n2_$mode.err").read()[-2000:])
PY
done
