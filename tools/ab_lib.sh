# usage: bash tools/ab_lib.sh LIB.so [LIB.so ...] -- bench the default line with each library swapped in for libsvo_b200.so (A/B of kernel builds)
mkdir -p gpurun_out
cp stereo-visual-odometry_b200/libsvo_b200.so /tmp/libsvo_b200_base.so
for lib in /tmp/libsvo_b200_base.so "$@"; do
  cp $lib stereo-visual-odometry_b200/libsvo_b200.so
  python bench.py --no-cpu-baseline --steps ${STEPS:-30} --warmup 5 2>gpurun_out/ab.err | tail -1 > gpurun_out/ab_last.json
  python -c "
import json,sys; d=json.load(open('gpurun_out/ab_last.json')); s=d['stage_ms_per_step']; print('$lib', 'step', round(d['ms_per_step'],4), 'lk', s['lk_track1'], 'value', round(d['value'],1))"
done
cp /tmp/libsvo_b200_base.so stereo-visual-odometry_b200/libsvo_b200.so
