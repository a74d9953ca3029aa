# quick GPU check: parity tests, then a short bench summary
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -${TAIL:-6}
python bench.py --no-cpu-baseline ${BENCH_ARGS:-} 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_last.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_last.json'))
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'])
print(d['stage_ms_per_step'])
print(d['clocks'])
PY
tail -3 gpurun_out/bench.err
