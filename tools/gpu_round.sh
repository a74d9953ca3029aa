# usage: TAG=r02c bash tools/gpu_round.sh -- tests, the default bench line (with the CPU baseline), the other configs, launch list, full ncu capture of one step
mkdir -p gpurun_out
TAG=${TAG:-r02}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 700 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
for cfg in ${CONFIGS:-c3 c4 c5}; do
  python bench.py --config $cfg ${CFG_ARGS:-} > gpurun_out/${TAG}_bench_${cfg}.json 2> gpurun_out/${TAG}_bench_${cfg}.err; tail -c 400 gpurun_out/${TAG}_bench_${cfg}.json; tail -2 gpurun_out/${TAG}_bench_${cfg}.err
done
# the pose step: bench line with K13 in the step (CPU arm with the real cv2.solvePnPRansac), and K13 alone on scenes with depth
python bench.py --pnp > gpurun_out/${TAG}_bench_pnp.json 2> gpurun_out/${TAG}_bench_pnp.err; tail -c 300 gpurun_out/${TAG}_bench_pnp.json; tail -2 gpurun_out/${TAG}_bench_pnp.err
python tools/bench_pnp.py > gpurun_out/${TAG}_pnp_alone.txt 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_pnp_alone_launches.csv python tools/bench_pnp.py > /dev/null 2>&1
cat gpurun_out/${TAG}_pnp_alone.txt
if [ -z "${NO_NCU:-}" ]; then
python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_l.log 2>&1
# one full step: 21 kernels per step (repitch, pyramid_fused, fast_cells, compact_scan, compact_gather, octree, describe, match_tc, orb_filter, fast_full, compact_full, 4 lk_pad + 3 pyr_down, lk_track (whole chain), lk_filter, accum_flags)
python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 63 -c 21 -f -o gpurun_out/${TAG}_full python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-200
fi
ls -la gpurun_out | tail -8
