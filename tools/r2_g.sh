mkdir -p gpurun_out/r2g
python -m pytest tests/test_gpu_cli.py tests/test_gpu_batch.py -x -q 2>&1 | tail -3
for k in hzcc_pack_kernel hzcc_scan_kernel hzdec_token_kernel sbt_inv_tile_kernel sbt_fwd_tile_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 -o gpurun_out/r2g/ncu_I_$k -f python tools/ab_kernel.py 64 > gpurun_out/r2g/ncu_I_$k.log 2>&1
done
