ADMP_SCF_HOSTSYNC=1 python tools/run_config.py C3 1 > gpurun_out/c3_plain.log 2>&1 && \
ADMP_SCF_HOSTSYNC=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 260 --csv --log-file gpurun_out/launches_c3cycle.csv python tools/run_config.py C3 1 > gpurun_out/ncu_c3cycle.log 2>&1
python tools/launch_summary.py gpurun_out/launches_c3cycle.csv "C3 SCF cycles (hostsync), launches 300-560" | head -30 | cut -c1-150
