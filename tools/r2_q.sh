python -m pytest tests/test_gpu_motion.py -x -q 2>&1 | tail -3
for v in base "" ing1 ing2 ing8; do echo "== ${v:-new}"; if [ -n "$v" ]; then export DSV1_B200_LIB=digital-subband-video-1_b200/build/ab/libdsv1_b200_$v.so; else unset DSV1_B200_LIB; fi; python tools/flag_probe.py 2>&1 | grep -E "hme_l0|ingest|hme_level"; done
