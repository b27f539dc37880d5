mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:col_direct -s 300 -c 1 -o gpurun_out/col_r1b python prof5_tmp.py > gpurun_out/ncu_col.log 2>&1
tail -2 gpurun_out/ncu_col.log
