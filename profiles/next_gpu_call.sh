#!/bin/bash
# First GPU call of the next round: everything written after round 1's GPU budget ran out, each step under its own
# timeout (a hung step must not eat the budget: round 1 lost 100 GPU-minutes to a 2-rank run that did not exit).
#   gpurun --timeout 900 -- 'bash profiles/next_gpu_call.sh'
mkdir -p gpurun_out
echo "== edge-case tests"; timeout 300 python -m pytest tests/test_zz_gpu_edge_cases.py -q 2>&1 | tail -5
echo "== glue tests";      LFGC_TEST_GLUE=1 timeout 300 python -m pytest tests/test_gpu_trainer.py -q -k glue 2>&1 | tail -8
echo "== bench, separate kernels"; timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/next_base.json 2> gpurun_out/next_base.err
echo "== bench, glue";             LFGC_GLUE=1 timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/next_glue.json 2> gpurun_out/next_glue.err
python - <<'PY'
import json
for tag in ('base', 'glue'):
    try:
        d = json.loads([l for l in open('gpurun_out/next_%s.json' % tag) if l.startswith('{')][-1])
        print(tag, 'samples/s %.4g' % d['value'], 'us/step %.2f' % d['extra']['us_per_optimiser_step'],
              'launches/step', d['extra']['launches_per_optimiser_step'], 'e2e %.4g' % d['e2e']['value'])
    except Exception as e:
        print(tag, 'no result:', e)
PY
echo "== tcgen05 kernel with dedicated MMA warp (LFGC_TC_ISSUER=1): parity + time against the default"
LFGC_TC_ISSUER=1 timeout 200 python profiles/tc_train_check.py 2>&1 | grep "TC=\|fused\|backward-only" | head -16
LFGC_TC_ISSUER=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -q 2>&1 | tail -3
LFGC_TC_ISSUER=1 timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('issuer-warp: us/step %.2f kernel_us %.2f' % (d['extra']['us_per_optimiser_step'], d['roofline']['kernel_us']))"
