#!/bin/bash
N=2
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu $EXTRA > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err; python -c "
import json
d=json.load(open('gpurun_out/bench_n${N}_$tag.json'))
print('N=$N $tag:', round(d['value'],1), 'ex/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1), 'replica', d['replica_check']['max_abs_param_diff_vs_rank0'], d['config']['gradient_exchange'][:36], 'loss', d['final_loss'])" || tail -5 gpurun_out/bench_n${N}_$tag.err; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py > gpurun_out/ddp_check_r2u.txt 2>&1; echo "ddp_check rc $?"; grep "multicast\|ddp_check ok" gpurun_out/ddp_check_r2u.txt || tail -20 gpurun_out/ddp_check_r2u.txt
EXTRA="" run mc B200MART_MULTICAST=1
EXTRA="" run mc_t64 B200MART_MULTICAST=1 B200MART_MC_TAIL_CTAS=64
EXTRA="" run mc_t32 B200MART_MULTICAST=1 B200MART_MC_TAIL_CTAS=32
EXTRA="--no-multicast" run nccl_ar X=0
