python -m pytest tests/test_gpu_fft.py tests/test_gpu_parallel.py -x -q 2>&1 | tail -2
echo "== C2"; python tools/xpass_time.py 1 1 1 2>&1 | grep -E "fft_z|fused"
echo "== C3"; python tools/xpass_time.py 2 4 4 2>&1 | grep -E "fft_z|fused"
echo "== C5"; python tools/xpass_time.py 4 8 8 2>&1 | grep -E "fft_z|fused"
