# Scaling runs on ONE box (gpurun --gpus 8): the driver's command line for N = 1, 2, 4, 8, stack config (with and
# without the layer-wise overlapped all-reduce at N = 8) and the full-model config at N = 1 and 8.
# COST: an 8-GPU box is charged 8x its wall time, and every default bench run also times the stock-PyTorch and CPU
# baselines on rank 0 (~1 min each): the round-2 call took 17 min of wall time = the whole remaining GPU budget, and
# the last line (whole model, N = 8) was cut off.  Pass `stack` or `full` to run only one group per call.
GROUP=${1:-all}
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$GROUP" != full ]; then
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
  echo "N=$n rc=$?"
done
$TR --nproc-per-node 8 --master-port 29650 bench.py --gpus 8 --steps 10 --warmup 3 --no-overlap-allreduce > gpurun_out/r02_scale_n8_single_allreduce.json 2> gpurun_out/r02_scale_n8_single_allreduce.err
echo "N=8 single rc=$?"
fi
if [ "$GROUP" != stack ]; then
python bench.py --config full --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_scale_full_n1.json 2> gpurun_out/r02_scale_full_n1.err
$TR --nproc-per-node 8 --master-port 29660 bench.py --config full --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_scale_full_n8.json 2> gpurun_out/r02_scale_full_n8.err
echo "full N=8 rc=$?"
fi
python - <<'PY'
import json
for f in ["n1","n2","n4","n8","n8_single_allreduce","full_n1","full_n8"]:
    try:
        d=json.load(open(f"gpurun_out/r02_scale_{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1), d["engine"].get("allreduce"))
    except Exception as e: print(f, "ERR", e)
PY
