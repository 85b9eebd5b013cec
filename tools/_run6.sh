python -m pytest tests/test_gpu_fft.py -x -q 2>&1 | tail -2
for cfg in "" "ADMP_FFT_XWIDE=0" "ADMP_FFT_XCFG=7"; do
echo "== C3 $cfg"; env $cfg python tools/xpass_time.py 2 4 4 2>&1 | grep -E "fft_x|fft_y_fwd|fused"
done
for cfg in "" "ADMP_FFT_XWIDE=0" "ADMP_FFT_XCFG=8"; do
echo "== C5 $cfg"; env $cfg python tools/xpass_time.py 4 8 8 2>&1 | grep -E "fft_x|fused"
done
echo "== C2"; python tools/xpass_time.py 1 1 1 2>&1 | grep -E "fft_|fused"
