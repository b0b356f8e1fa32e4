python tools/flag_probe.py 2>&1 | grep -E "hme_l"
for v in hmA hmB hmC; do echo "== $v"; DSV1_B200_LIB=digital-subband-video-1_b200/build/ab/libdsv1_b200_$v.so python tools/flag_probe.py 2>&1 | grep -E "hme_l"; done
