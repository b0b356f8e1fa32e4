O=gpurun_out/r2t; mkdir -p $O
timeout 40 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu > $O/bench_config3.json 2> $O/bench_config3.err; python tools/parse_bench.py < $O/bench_config3.json | head -1
