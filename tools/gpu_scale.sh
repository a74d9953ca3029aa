# usage (under gpurun --gpus N): N=4 TAG=r02g bash tools/gpu_scale.sh -- c1 / c3 / c5 bench lines at N GPUs (torchrun), one JSON line each
mkdir -p gpurun_out
N=${N:-2}
TAG=${TAG:-r02}
P=29700
for cfg in ${CONFIGS:-c1 c3 c5}; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --config $cfg --no-cpu-baseline 2> gpurun_out/${TAG}_${cfg}_n${N}.err | grep "^{" | tail -1 > gpurun_out/${TAG}_bench_${cfg}_n${N}.json
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${cfg}_n${N}.json"))
print("${cfg}", "N=${N}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), d.get("merged_signature"))
PY
done
