#!/bin/bash
# A/B runs of the default bench under environment toggles: tools/ab_env.sh "VAR=1" "VAR=2 OTHER=3" ...
run() { echo "== $*"; env "$@" python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'ex/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1), ' loss', round(d['final_loss'],4))"; }
for a in "$@"; do run $a; done
