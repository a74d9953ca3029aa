mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_orb_gpu.py tests/test_lk_gpu.py tests/test_frontend_gpu.py -m gpu -x -q -k "640 or 333 or border or levels or flat or match_hamming or reset" > gpurun_out/sanitize.log 2>&1; echo "sanitizer rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize.log | head -20
