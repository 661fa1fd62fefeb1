#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kernel_ab.py --variants notos,base,ctas8,ctas9,ctas10 --json gpurun_out/r2b_ab.json horse_and_mug:3840:1920:16 horse_and_mug:1440:720:1 marbles:2048:2048:4 car:2048:1536:8 > gpurun_out/r2b_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/r2b_ab.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2b_bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','render_kernel_ms')}, d['e2e'], d['scene_build'])
    for c in d.get('configs',[]): print(c)
    print(d.get('cli_wall'))
    print(d.get('cpu_baseline'))
except Exception as e: print('bench parse failed', e)
PY
