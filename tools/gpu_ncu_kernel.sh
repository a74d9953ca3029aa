# usage: KREGEX=k_lk_track SKIP=8 COUNT=4 NAME=prof bash tools/gpu_ncu_kernel.sh
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --steps 3 --warmup 3 --frames 16 > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -s ${SKIP:-8} -c ${COUNT:-4} -o gpurun_out/${NAME:-prof} -f python bench.py --no-cpu-baseline --steps 3 --warmup 3 --frames 16 > gpurun_out/ncu_kernel.log 2>&1
tail -2 gpurun_out/ncu_kernel.log | cut -c1-200
