set -x
PROF=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'radix_hist_kernel|radix_scatter_kernel|boundaries_kernel|diag_' -c 6 -o gpurun_out/prof_conv3_r01 -f python scripts/prof_c3c4.py uniform > gpurun_out/conv3_ncu.log 2>&1
tail -2 gpurun_out/conv3_ncu.log
