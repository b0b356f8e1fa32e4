mkdir -p gpurun_out/r2f
python tools/diag_contend.py 2>&1 | head -3
python tools/diag_e2e.py 64
python tools/timeline_e2e.py 2>/dev/null | head -8
python bench.py > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err; tail -3 gpurun_out/r2f/bench.err; python tools/parse_bench.py < gpurun_out/r2f/bench.json | head -40
