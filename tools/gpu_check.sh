#!/bin/bash
# full GPU check: every -m gpu test, then the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -120 > gpurun_out/check_tests.txt
tail -4 gpurun_out/check_tests.txt; grep "^E  \|^FAILED" gpurun_out/check_tests.txt | head -10
timeout 600 python bench.py --steps 30 --warmup 5 ${BENCH_ARGS} > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/check_bench.json'))
    print('ms_per_step', round(d['ms_per_step'],4), 'fps', round(d['value']), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None, 'launches/step', d['gpu_launches_per_step'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks'].get('samples'))
    print('roofline frac', round(d['roofline']['frac'],3), [ (r['kernel'], round(r.get('frac',0),3)) for r in d['roofline_kernels']])
    print({k:v['ms_per_step'] for k,v in d['breakdown_eager_ms'].items()})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/check_bench.err').read()[-2000:])
PY
