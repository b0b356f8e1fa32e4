python -m pytest tests/test_gpu_hzcc_enc.py -x -q 2>&1 | tail -1
python tools/flag_probe.py 2>&1 | grep -E "hzcc_"
for v in hzb8 hzb2; do echo "== $v"; DSV1_B200_LIB=digital-subband-video-1_b200/build/ab/libdsv1_b200_$v.so python tools/flag_probe.py 2>&1 | grep -E "hzcc_scan|hzcc_pack_k"; done
