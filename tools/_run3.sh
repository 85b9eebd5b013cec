PROF_ONCE=1 python tools/prof_recip.py 2 4 4 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fast_x_conv|fast_strided|fast_z' -f -o /tmp/x_244 env PROF_ONCE=1 python tools/prof_recip.py 2 4 4 1 > gpurun_out/ncu_x_244.log 2>&1
ncu -i /tmp/x_244.ncu-rep --page details > gpurun_out/x_244.details.txt 2>/dev/null
ncu -i /tmp/x_244.ncu-rep --page source --csv > gpurun_out/x_244.source.csv 2>/dev/null
ncu -i /tmp/x_244.ncu-rep --page raw --csv > gpurun_out/x_244.raw.csv 2>/dev/null
ls -la gpurun_out/x_244*
