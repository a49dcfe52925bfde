#!/bin/bash
# Round 2, call 12: K-chunked FiLM head (no 403 MB outer-product matrix): parity, bench of the film variant.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -s -k "film" > gpurun_out/r2c12_tests.log 2>&1
echo "== film tests exit $?"; grep -E "passed|failed|^FAILED|Error|^E |forced film" gpurun_out/r2c12_tests.log | tail -14 | cut -c1-400
for f in film concat; do
  timeout 300 python bench.py --fusion $f --steps 10 --warmup 3 --no-cpu --no-device-pipeline > gpurun_out/r2c12_bench_$f.log 2>&1
  echo "== bench $f exit $?"; grep '^{"metric"' gpurun_out/r2c12_bench_$f.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
print('   ms/step %.3f value %.0f e2e %.0f launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['launches_per_step']))
print('   ' + ' '.join('%s=%.2f' % (k, v['ms']) for k, v in kb.items()))
"
done
