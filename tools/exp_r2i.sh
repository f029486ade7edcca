#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py > gpurun_out/ddp_check_r2i.txt 2>&1; echo "ddp_check rc $?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/ddp_check_r2i.txt | tail -6
run() { tag=$1; shift; timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu $EXTRA > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err; python -c "
import json
d=json.load(open('gpurun_out/bench_n2_$tag.json'))
print('$tag:', round(d['value'],1), 'ex/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1), 'replica', d['replica_check'], d['config']['gradient_exchange'][:40])" || tail -5 gpurun_out/bench_n2_$tag.err; }
EXTRA="" run mc X=0
EXTRA="--no-multicast --no-shard" run nccl_ar X=0
EXTRA="" run mc2 X=0
