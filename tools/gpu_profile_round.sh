# usage: TAG=r01b bash tools/gpu_profile_round.sh  -- bench line, launch list and one full capture of every kernel of one step
mkdir -p gpurun_out
TAG=${TAG:-r01}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json
python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_l.log 2>&1
# one full step: 16 kernels per step (repitch, pyramid_fused, fast_cells, compact, octree, describe, match, orb_filter, fast_full, compact_full, 3 pyr_down, lk_pad, lk_track (whole chain), lk_filter)
python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 48 -c 16 -f -o gpurun_out/${TAG}_full python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-200
ls -la gpurun_out | tail -8
