# A/B of the LK chain launch's grid (one-warp CTAs per frame): SVO_LK_GRID in {pt_cap (old behaviour), 4096, 2048, 1024, 512}
mkdir -p gpurun_out
python -m pytest tests/test_lk_gpu.py tests/test_frontend_gpu.py -x -q 2>&1 | tail -2
for g in 100000 4096 2048 1024 768 512; do
  SVO_LK_GRID=$g python bench.py --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('grid $g', 'ms/step', round(d['ms_per_step'],3), 'lk', d['stage_ms_per_step']['lk_track1'], 'value', round(d['value'],1))"
done
