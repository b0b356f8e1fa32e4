python -m pytest tests/test_gpu_sbt.py tests/test_gpu_stream.py -x -q 2>&1 | tail -3
python tools/flag_probe.py 2>&1 | grep -E "sbt_|hzcc_scan|clean"
for v in minb5 minb4; do echo "== $v"; DSV1_B200_LIB=digital-subband-video-1_b200/build/ab/libdsv1_b200_$v.so python tools/flag_probe.py 2>&1 | grep -E "sbt_inv_tile"; done
