#!/bin/bash
# first GPU run of the round-2 rewrite: smoke, the GPU test-suite (every failure listed), creation profile, kernel A/B
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
tail -5 gpurun_out/r2a_smoke.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2a_pytest.log
timeout 300 python tools/scene_create_profile.py > gpurun_out/r2a_create.jsonl 2> gpurun_out/r2a_create.err; echo "create rc=$?"
cat gpurun_out/r2a_create.jsonl | cut -c1-400
timeout 900 python tools/kernel_ab.py --variants base,nodiv3,accshared,ctas8,ctas6 --json gpurun_out/r2a_ab.json > gpurun_out/r2a_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/r2a_ab.log
