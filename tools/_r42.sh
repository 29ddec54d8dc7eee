mkdir -p gpurun_out
timeout 300 python tools/tc_debug.py > gpurun_out/r42_tcdebug.log 2>&1
echo "tc_debug ok lines: $(grep -c 'timeouts 0' gpurun_out/r42_tcdebug.log)"; grep -v "timeouts 0" gpurun_out/r42_tcdebug.log | head -5
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r42_pytest.log 2>&1; tail -2 gpurun_out/r42_pytest.log
for s in conv1y conv2y refine0_upconv pd0_conv1 conv1x; do TC_TIMING=1 timeout 120 python tools/bench_conv.py $s 1 3; done > gpurun_out/r42_timing.log 2>&1
grep -E "done|MMA total|epilogue|accum" gpurun_out/r42_timing.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r42_bench.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r42_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['instrumented_ms_per_step'])
        for k in d['roofline']['kernels']: print('   ', k['kernel'][:50], round(k['ms_per_step'],2), round(k['tflops'],1))
PY
