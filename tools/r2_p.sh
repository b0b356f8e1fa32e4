python tools/flag_probe.py 2>&1 | grep -E "bmc"
for v in rl48 rl16 rl48d; do echo "== $v"; DSV1_B200_LIB=digital-subband-video-1_b200/build/ab/libdsv1_b200_$v.so python tools/flag_probe.py 2>&1 | grep -E "bmc"; done
