python -m pytest tests/test_gpu_hzcc_enc.py tests/test_gpu_stream.py -x -q 2>&1 | tail -2
python tools/flag_probe.py 2>&1 | grep -E "hzcc_"
