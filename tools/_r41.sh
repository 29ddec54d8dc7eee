mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo -c 1 -o gpurun_out/r41_conv1y -f python tools/bench_conv.py conv1y 1 1 > gpurun_out/r41_ncu.log 2>&1
tail -2 gpurun_out/r41_ncu.log
