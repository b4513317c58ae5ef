#!/bin/bash
# Two-GPU checks of the next round, each under its own timeout:
#   gpurun --gpus 2 --timeout 600 -- 'bash profiles/next_gpu_call_2gpu.sh'
mkdir -p gpurun_out
run() {  # tag, extra env
    env $2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline \
        > gpurun_out/next2_$1.json 2> gpurun_out/next2_$1.err
    echo "$1: exit $? after ${SECONDS}s"
    python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open('gpurun_out/next2_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], 'samples/s %.4g' % d['value'], 'us/step %.2f' % d['extra']['us_per_optimiser_step'])
except Exception as e:
    print(sys.argv[1], 'no result:', e)
PY
}
run nccl "LFGC_ALLREDUCE=nccl"          # also checks that the run EXITS now (round 1: it did not)
run symm "LFGC_ALLREDUCE=symm"          # one-shot all-reduce over symmetric memory
run symm_glue "LFGC_ALLREDUCE=symm LFGC_GLUE=1"
run p2p "LFGC_ALLREDUCE=p2p"            # all-reduce folded into the Adam kernel (lfgc_adam_p2p) between symm-mem barriers
