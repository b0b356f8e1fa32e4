for v in base new; do echo "== $v"; if [ $v = base ]; then export DSV1_B200_LIB=digital-subband-video-1_b200/build/ab/libdsv1_b200_base.so; else unset DSV1_B200_LIB; fi; python tools/flag_probe.py 2>&1 | grep -E "hzcc_|pack_kernel"; done
python -m pytest tests/test_gpu_hzcc_enc.py tests/test_gpu_stream.py -x -q 2>&1 | tail -2
