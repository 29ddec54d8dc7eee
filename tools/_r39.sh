mkdir -p gpurun_out
for s in conv1y conv2y; do TC_TIMING=1 timeout 120 python tools/bench_conv.py $s 1 3; done > gpurun_out/r39_timing.log 2>&1
cat gpurun_out/r39_timing.log
