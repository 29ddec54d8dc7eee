mkdir -p gpurun_out
timeout 300 python tools/tc_debug.py > gpurun_out/r34_tcdebug.log 2>&1
grep -c "timeouts 0" gpurun_out/r34_tcdebug.log; grep -v "timeouts 0" gpurun_out/r34_tcdebug.log | head
for s in pd0_conv1 conv1x refine0_upconv predict2_conv1 refine_conv1_1; do TC_TIMING=1 timeout 120 python tools/bench_conv.py $s 1 3; done > gpurun_out/r34_timing.log 2>&1
grep -E "done|MMA total|MMA wait T_full|stager0 wait T_empty" gpurun_out/r34_timing.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r34_pytest.log 2>&1
tail -2 gpurun_out/r34_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r34_bench.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r34_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'])
PY
