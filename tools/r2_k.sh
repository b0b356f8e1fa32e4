python -m pytest tests/test_gpu_sbt.py tests/test_gpu_hzcc_enc.py -x -q 2>&1 | tail -2
python tools/flag_probe.py 2>&1 | grep -E "sbt_fwd_tile|sbt_inv_tile|clean"
python tools/quick_time.py hd_gop0 hd_gop12 2>&1 | grep -v "^ref"
python tools/quick_time.py cif_gop12 uhd444_gop12 2>&1 | grep -v "^ref"
