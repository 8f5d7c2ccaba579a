#!/bin/bash
# usage: tools/run_multi_gpu.sh N [bench args...]   (one node, N ranks, NCCL)
N=$1; shift
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$N" "$@"
