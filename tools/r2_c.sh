mkdir -p gpurun_out/r2c
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2c/gputest.log
cat gpurun_out/r2c/gputest.log
python tools/ab_kernel.py 64 > gpurun_out/r2c/ab.log 2>&1; cat gpurun_out/r2c/ab.log
python bench.py > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err; head -c 1200 gpurun_out/r2c/bench.json
