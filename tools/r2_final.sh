# round-2 closing evidence: tests, bench lines, launch list, ncu --set full summaries, timeline, sanitizers
O=gpurun_out/r2z; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/gputest.txt; cat $O/gputest.txt
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; python tools/parse_bench.py < $O/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench_n1.err
for c in 2 3 4; do python bench.py --config $c --steps 3 --warmup 3 > $O/bench_config$c.json 2> $O/bench_config$c.err; python tools/parse_bench.py < $O/bench_config$c.json | head -2; done
python tools/quick_time.py cif_gop12 hd_gop0 hd_gop12 uhd444_gop12 > $O/sync_api.txt 2>&1; cat $O/sync_api.txt
python tools/timeline_e2e.py --steps 6 > $O/timeline_e2e.txt 2>/dev/null; head -7 $O/timeline_e2e.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/bench_under_ncu.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launches.txt 2>&1; rm -f $O/launches.csv; head -12 $O/launches.txt
for k in sbt_inv_tile_kernel sbt_fwd_tile_kernel bmc_kernel hme_l0_kernel hzcc_scan_kernel hzcc_pack_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"$k\$" -s 4 -c 2 -o $O/ncu_$k -f python tools/ab_kernel.py 64 > $O/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:sbt_inv_tile_intra_kernel -s 0 -c 1 -o $O/ncu_sbt_inv_tile_intra_kernel -f python tools/ab_kernel.py 64 > $O/ncu_intra.log 2>&1
python tools/ncu_traffic.py $O/traffic.json $O/ncu_*.ncu-rep > /dev/null
for f in $O/ncu_*.ncu-rep; do b=$(basename $f .ncu-rep); python tools/ncu_summary.py $f > $O/$b.txt; python tools/op_hist.py $f "${b#ncu_}" > $O/${b}_ops.txt 2>/dev/null; done
for k in sbt_inv_tile_kernel sbt_fwd_tile_kernel; do python tools/src_hot.py $O/ncu_$k.ncu-rep $k 40 > $O/ncu_${k}_lines.txt 2>/dev/null; done
rm -f $O/ncu_bmc_kernel.ncu-rep $O/ncu_hme_l0_kernel.ncu-rep $O/ncu_hzcc_scan_kernel.ncu-rep $O/ncu_hzcc_pack_kernel.ncu-rep $O/ncu_sbt_fwd_tile_kernel.ncu-rep $O/ncu_*.log
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_run.py > $O/sanitizer_$tool.txt 2>&1; tail -1 $O/sanitizer_$tool.txt
done
cuobjdump -sass digital-subband-video-1_b200/libdsv1_b200.so | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{print $2}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn > $O/sass_opcodes.txt; grep -cE "UTMA|TCGEN|UBLKCP" $O/sass_opcodes.txt
du -sh $O
