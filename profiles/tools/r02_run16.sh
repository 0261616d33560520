# 2 GPUs: the slab form of the host input (row slabs over PCIe + column exchange over NVLink) -- parity on hardware and N=2 e2e
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02p_build.log 2>&1
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -v > gpurun_out/r02p_pytest_sharded_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02p_pytest_sharded_2gpu.log
tail -6 gpurun_out/r02p_pytest_sharded_2gpu.log
run() {
  env $3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $2 --steps 5 --warmup 3 $4 > gpurun_out/r02p_$1.json 2> gpurun_out/r02p_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02p_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["ms_per_step"],3), d.get("e2e"), d["self_check"]["root_match"], d["result"]["root"][:16])
except Exception as e:
    print("$1 ERR", e); print(open("gpurun_out/r02p_$1.err").read()[-1500:])
PY
}
run n2_slab 2 TS_HOST_INPUT=slab ""
run n2_window 2 TS_HOST_INPUT=window ""
