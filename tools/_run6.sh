echo "== C3 WIDE=1"; ADMP_FFT_WIDE=1 python tools/xpass_time.py 2 4 4 2>&1 | grep -E "fft_z|fused"
echo "== C5 WIDE=1"; ADMP_FFT_WIDE=1 python tools/xpass_time.py 4 8 8 2>&1 | grep -E "fft_z|fused"
echo "== C2 WIDE=1"; ADMP_FFT_WIDE=1 python tools/xpass_time.py 1 1 1 2>&1 | grep -E "fft_z|fused"
