mkdir -p gpurun_out
set -x
python bench.py --no-cpu-baseline > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -c 2500 gpurun_out/bench_r1a.json; tail -3 gpurun_out/bench_r1a.err
python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log; wc -l gpurun_out/launches_r1a.csv
