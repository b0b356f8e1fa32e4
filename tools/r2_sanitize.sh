set -x
mkdir -p gpurun_out/r2b
python -m pytest tests/test_gpu_long.py -x -q 2>&1 | tail -15 > gpurun_out/r2b/gpu_long.log
cat gpurun_out/r2b/gpu_long.log
python tools/sanitize_run.py > gpurun_out/r2b/plain.log 2>&1; tail -2 gpurun_out/r2b/plain.log
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_run.py > gpurun_out/r2b/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/r2b/sanitizer_$tool.log
done
