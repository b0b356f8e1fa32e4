for v in base new b6; do echo "== $v"; DSV1_B200_LIB=digital-subband-video-1_b200/build/ab/libdsv1_b200_$v.so python tools/flag_probe.py 2>&1 | grep -E "hme_l0|ingest|down2"; done
python -m pytest tests/test_gpu_motion.py -x -q 2>&1 | tail -2
