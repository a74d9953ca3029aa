mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline --steps 3 --warmup 3 --frames 16 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_lk_track|k_fast_cells|k_describe' -s 12 -c 6 -o gpurun_out/prof_r1a python bench.py --no-cpu-baseline --steps 3 --warmup 3 --frames 16 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log | cut -c1-300
ls -la gpurun_out/
