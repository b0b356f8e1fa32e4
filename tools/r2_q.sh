O=gpurun_out/r2r; mkdir -p $O
python tools/quick_time.py hd_gop12 cif_gop12 2>&1 | grep -v "^ref" | tail -4
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; python tools/parse_bench.py < $O/bench_n1.json | head -40
