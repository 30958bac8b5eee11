#!/bin/bash
# gpurun --gpus N calls: C3 weak scaling at 8 GPUs (NCCL exchange vs peer-memory exchange), BASELINE configs[2..4]
# (8192 hexamer walkers; C4 4096; C5 2048) and the strong-scaling points of configs[2] (8192 walkers over 2 / 4 GPUs).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name nproc extra-env args...
    local name=$1 n=$2 envs=$3; shift 3
    env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus "$n" --steps 20 --warmup 5 --no-cpu --no-dropin "$@" > "gpurun_out/$name.json" 2> "gpurun_out/$name.err"
    echo "$name rc=$? $(head -c 300 gpurun_out/$name.json)"
}
# what to run: "$1" = n8 (default), peer, n4, n2 -- one gpurun call each, so that a call only pays for the GPUs it uses
case "${1:-n8}" in
n8)   run r02g_bench_c3_n8_nccl 8 HANS_NS_PEER=0 --others C4,C5 ;;
peer) run r02g_bench_c3_n8_peer 8 HANS_NS_PEER=1 --others "" ;;
n4)   run r02g_bench_c3_strong_n4 4 HANS_NS_PEER=0 --total-walkers 8192 --others "" ;;
n2)   run r02g_bench_c3_strong_n2 2 HANS_NS_PEER=0 --total-walkers 8192 --others "" ;;
esac
