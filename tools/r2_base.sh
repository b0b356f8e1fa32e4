set -x
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2a/gputest.log
python bench.py > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a/bench_ref.json 2>> gpurun_out/r2a/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2a/launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r2a/bench_under_ncu.log 2>&1
for k in sbt_inv_tile_kernel bmc_kernel hme_l0_kernel hzcc_scan_kernel hzcc_pack_kernel sbt_fwd_tile_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -o gpurun_out/r2a/ncu_$k -f python tools/ab_kernel.py 64 > gpurun_out/r2a/ncu_$k.log 2>&1
done
tail -3 gpurun_out/r2a/gputest.log; head -c 1500 gpurun_out/r2a/bench.json
