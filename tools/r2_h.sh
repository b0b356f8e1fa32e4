mkdir -p gpurun_out/r2h
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2h/gputest.log; cat gpurun_out/r2h/gputest.log
python tools/quick_time.py cif_gop12 hd_gop0 hd_gop12 uhd444_gop12 2>&1 | tee gpurun_out/r2h/sync_api.txt
for c in 2 3 4; do python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r2h/bench_config$c.json 2> gpurun_out/r2h/bench_config$c.err; python tools/parse_bench.py < gpurun_out/r2h/bench_config$c.json | head -2; done
# one whole 12-picture pass (1 I + 11 P launches of every kernel) under ncu --set full: traffic per launch as the bench averages it
ncu --set full --clock-control none --import-source on -k regex:'sbt_inv_tile_kernel|sbt_fwd_tile_kernel|bmc_kernel|hme_l0_kernel|hzcc_scan_kernel|hzcc_pack_kernel' -s 0 -c 80 -o gpurun_out/r2h/ncu_pass_b64 -f python tools/ab_kernel.py 64 > gpurun_out/r2h/ncu_pass.log 2>&1
tail -3 gpurun_out/r2h/ncu_pass.log
