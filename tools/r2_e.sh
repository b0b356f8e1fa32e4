mkdir -p gpurun_out/r2e
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2e/gputest.log
cat gpurun_out/r2e/gputest.log
python tools/flag_probe.py > gpurun_out/r2e/probe.log 2>&1; cat gpurun_out/r2e/probe.log
