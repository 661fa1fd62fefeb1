#!/bin/bash
# bench.py at N GPUs (both gather modes): usage tools/r2_scale.sh N [tag]
N=${1:-2}; T=${2:-r2}
mkdir -p gpurun_out
for G in nccl p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --gather $G > gpurun_out/${T}_bench_n${N}_$G.json 2> gpurun_out/${T}_bench_n${N}_$G.err; echo "bench $G rc=$?"
tail -1 gpurun_out/${T}_bench_n${N}_$G.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('n_gpus','value','ms_per_step','render_kernel_ms')}, 'e2e', d['e2e']['ms_per_frame'], d['render_kernel_ms_per_rank'], d['frame']['frame_sha256'][:12], d['frame']['device_frame_sha256'][:12], d['clocks'])"
done
