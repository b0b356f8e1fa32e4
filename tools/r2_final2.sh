O=gpurun_out/r2y; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/gputest.txt; cat $O/gputest.txt
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; python tools/parse_bench.py < $O/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench_n1.err
for c in 2 3 4; do python bench.py --config $c --steps 3 --warmup 3 > $O/bench_config$c.json 2> $O/bench_config$c.err; python tools/parse_bench.py < $O/bench_config$c.json | head -1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/bench_under_ncu.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launches.txt 2>&1; rm -f $O/launches.csv
for k in sbt_inv_tile_kernel hzcc_scan_kernel hzcc_pack_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"$k\$" -s 4 -c 2 -o $O/ncu_$k -f python tools/ab_kernel.py 64 > $O/ncu_$k.log 2>&1
done
ncu --set full --clock-control none -k regex:hzcc_pack_dense_kernel -s 0 -c 1 -o $O/ncu_hzcc_pack_dense_kernel -f python tools/ab_kernel.py 64 > $O/ncu_dense.log 2>&1
python tools/ncu_traffic.py $O/traffic_update.json $O/ncu_*.ncu-rep > /dev/null
for f in $O/ncu_*.ncu-rep; do b=$(basename $f .ncu-rep); python tools/ncu_summary.py $f > $O/$b.txt; python tools/op_hist.py $f "${b#ncu_}" > $O/${b}_ops.txt 2>/dev/null; done
python tools/src_hot.py $O/ncu_sbt_inv_tile_kernel.ncu-rep sbt_inv_tile_kernel 40 > $O/ncu_sbt_inv_tile_kernel_lines.txt 2>/dev/null
rm -f $O/*.ncu-rep $O/ncu_*.log
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_run.py > $O/sanitizer_$tool.txt 2>&1; tail -1 $O/sanitizer_$tool.txt; done
du -sh $O
