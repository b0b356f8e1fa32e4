# round-2 closing evidence, part 2: ncu --set full of the kernels changed last, launch list, config 2, sanitizers
O=gpurun_out/r2t; mkdir -p $O
for k in hme_l0_kernel hzcc_scan_kernel hzcc_pack_kernel ingest_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"$k\$" -s 4 -c 2 -o $O/ncu_$k -f python tools/ab_kernel.py 64 > $O/ncu_$k.log 2>&1
done
python tools/ncu_traffic.py $O/traffic_update.json $O/ncu_*.ncu-rep > /dev/null
for f in $O/ncu_*.ncu-rep; do b=$(basename $f .ncu-rep); python tools/ncu_summary.py $f > $O/$b.txt; python tools/op_hist.py $f "${b#ncu_}" > $O/${b}_ops.txt 2>/dev/null; done
python tools/src_hot.py $O/ncu_hme_l0_kernel.ncu-rep hme_l0_kernel 40 > $O/ncu_hme_l0_kernel_lines.txt 2>/dev/null
rm -f $O/ncu_*.ncu-rep $O/ncu_*.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/bench_under_ncu.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launches.txt 2>&1; rm -f $O/launches.csv; head -8 $O/launches.txt
python bench.py --config 2 --steps 3 --warmup 3 > $O/bench_config2.json 2> $O/bench_config2.err; python tools/parse_bench.py < $O/bench_config2.json | head -1
for tool in memcheck racecheck; do timeout 200 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_run.py > $O/sanitizer_$tool.txt 2>&1; tail -1 $O/sanitizer_$tool.txt; done
du -sh $O
