#!/bin/bash
# the numbers profiles/ and DESIGN.md quote for one GPU (run via gpurun)
T=${1:-r2}
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_reference_n1.json 2> gpurun_out/${T}_reference_n1.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${T}_launches_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python tools/cli_wall.py --runs 4 --json gpurun_out/${T}_cli_wall.json --md gpurun_out/${T}_cli_wall.md > gpurun_out/${T}_cli_wall.log 2>&1; echo "cli rc=$?"
timeout 300 python tools/scene_create_profile.py > gpurun_out/${T}_create.jsonl 2>&1
if [ "$2" = "full" ]; then
timeout 1200 python bench.py --impl reference --full-frame > gpurun_out/${T}_reference_full_frame.json 2> gpurun_out/${T}_reference_full_frame.err; echo "full frame rc=$?"
fi
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','render_kernel_ms')}, d['e2e']['ms_per_frame'], d['scene_build']['cold_s'], d['scene_build']['warm_s'], d['clocks'])
r=json.loads(open('gpurun_out/${T}_reference_n1.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['steps'], r['warmup'], r['config']==d['config'])
PY
