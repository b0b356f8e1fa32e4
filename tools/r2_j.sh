python -m pytest tests/test_gpu_sbt.py tests/test_gpu_stream.py tests/test_gpu_long.py -x -q 2>&1 | tail -3
python tools/flag_probe.py 2>&1 | grep -E "sbt_|hzcc_scan|clean"
