python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== C5"; python tools/xpass_time.py 4 8 8 2>&1 | grep -E "fft_|fused"
echo "== C5 YCFG=6"; ADMP_FFT_YCFG=6 python tools/xpass_time.py 4 8 8 2>&1 | grep -E "fft_y|fused"
echo "== C3"; python tools/xpass_time.py 2 4 4 2>&1 | grep -E "fft_|fused"
