#!/bin/bash
# 2 GPUs: NCCL parity check (sharded vs replicated optimizer, replicas identical) + scaling A/B
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py > gpurun_out/ddp_check_r2h.txt 2>&1; echo "ddp_check rc $?"; tail -3 gpurun_out/ddp_check_r2h.txt
run() { tag=$1; shift; timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu $EXTRA > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err; python -c "
import json
d=json.load(open('gpurun_out/bench_n2_$tag.json'))
print('$tag:', round(d['value'],1), 'ex/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1), 'replica', d['replica_check'])" || tail -5 gpurun_out/bench_n2_$tag.err; }
EXTRA="" run shard X=0
EXTRA="--no-shard" run noshard X=0
EXTRA="" run shard_ctas8 NCCL_MAX_CTAS=8
EXTRA="" run shard_ctas16 NCCL_MAX_CTAS=16
EXTRA="" run shard_ctas4 NCCL_MAX_CTAS=4
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('n1:', round(d['value'],1), 'ex/s', round(d['ms_per_step'],3), 'ms')"
