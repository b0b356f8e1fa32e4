mkdir -p gpurun_out/r2d
for k in sbt_inv_tile_kernel hzcc_scan_kernel hzdec_clean_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -o gpurun_out/r2d/ncu_$k -f python tools/ab_kernel.py 64 > gpurun_out/r2d/ncu_$k.log 2>&1
done
