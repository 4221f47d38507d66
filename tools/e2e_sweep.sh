#!/bin/bash
# Diagnostic: repeated bench.py runs, printing value / e2e, the host phases, the drain and the per-block GPU periods of the e2e loop.
for a in "$@"; do
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-check-gather $a 2>gpurun_out/e.err > gpurun_out/e.json
  python - "$a" <<'PY'
import json, sys
d = json.load(open('gpurun_out/e.json'))
e = d['e2e']
print(sys.argv[1], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(e['value'], 1), round(e['ms_per_step'], 2), 'drain', round(e['drain_ms'], 1), e['host_phases'])
print('   gpu_block_ms', e['gpu_block_ms'])
PY
done
