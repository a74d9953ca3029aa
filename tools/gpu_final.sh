# usage: TAG=r02o bash tools/gpu_final.sh -- the short round-end capture: all GPU tests, the default bench line (with the CPU baseline), the c5 sweep
# (result signature), the Scharr-plane LK alternative on the LK tests, the launch list and one full ncu step
mkdir -p gpurun_out
TAG=${TAG:-r02}
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | grep -v "^$\|user\|sys"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
python bench.py --config c5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; tail -c 300 gpurun_out/${TAG}_bench_c5.json; tail -2 gpurun_out/${TAG}_bench_c5.err
SVO_LK_DERIV=1 python -m pytest tests/test_lk_gpu.py -m gpu -x -q 2>&1 | tail -1
python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -s 63 -c 21 -f -o gpurun_out/${TAG}_full python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-200
ls -la gpurun_out | tail -6
