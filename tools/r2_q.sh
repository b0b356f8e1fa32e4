for v in base pf d64 d320 pfd320; do echo "== $v"; DSV1_B200_LIB=digital-subband-video-1_b200/build/ab/libdsv1_b200_$v.so python tools/flag_probe.py 2>&1 | grep -E "hzcc_scan|hzcc_pack"; done
