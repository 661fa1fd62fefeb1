#!/bin/bash
# multi-GPU checks: usage tools/r2_multi_gpu.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG:-r2m}_smoke_n$N.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG:-r2m}_smoke_n$N.log
timeout 600 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -k "multi_gpu or bands or async" > gpurun_out/${TAG:-r2m}_pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG:-r2m}_pytest_n$N.log
for G in nccl p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --gather $G > gpurun_out/${TAG:-r2m}_bench_n${N}_$G.json 2> gpurun_out/${TAG:-r2m}_bench_n${N}_$G.err; echo "bench $G rc=$?"
tail -2 gpurun_out/${TAG:-r2m}_bench_n${N}_$G.err
tail -1 gpurun_out/${TAG:-r2m}_bench_n${N}_$G.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','value','ms_per_step','render_kernel_ms')}, d['e2e']['ms_per_frame'], d['frame']['frame_sha256'][:12], d['frame']['device_frame_sha256'][:12])"
done
python tools/cli_wall.py --runs 3 --gpus $N car:1 cornellbox:1 horse_and_mug:2 --json gpurun_out/${TAG:-r2m}_cli_n$N.json > gpurun_out/${TAG:-r2m}_cli_n$N.log 2>&1; echo "cli rc=$?"; tail -5 gpurun_out/${TAG:-r2m}_cli_n$N.log
