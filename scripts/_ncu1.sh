THSP_RADIX_VARIANT=4 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/conv_v4.csv python scripts/_one_conv.py > gpurun_out/conv_v4.log 2>&1
echo rc=$?
