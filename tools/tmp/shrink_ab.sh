python -m pytest tests/test_conv_tc_gpu.py tests/test_forward_gpu.py tests/test_forward_full_gpu.py -x -q 2>&1 | tail -2
python tools/fwd_profile.py 3xfp16 2>&1 | tail -2
CRESTE_TC_NO_SHRINK=1 python tools/fwd_profile.py 3xfp16 2>&1 | tail -2
