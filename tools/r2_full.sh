#!/bin/bash
# What the driver runs at round end: the whole GPU test suite, smoke(), the default bench, the reference arm.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 2400 python -m pytest tests/ -q -m gpu --no-header -p no:cacheprovider --durations=15 > gpurun_out/r2full_tests.log 2>&1
echo "== pytest -m gpu exit $? ($(( $(date +%s) - T0 )) s)"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2full_tests.log | tail -25 | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2full_smoke.log 2>&1
echo "== smoke exit $?"; tail -2 gpurun_out/r2full_smoke.log | cut -c1-300
GDL_DUMP_OPS=gpurun_out/r2full_ops.json timeout 600 python bench.py > gpurun_out/r2full_bench.log 2>&1
echo "== bench exit $?"; grep '^{"metric"' gpurun_out/r2full_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
print('   ms/step %.3f value %.0f e2e %.0f frac %.3f cpu %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value']))
print('   ' + ' '.join('%s=%.2f' % (k, v['ms']) for k, v in kb.items()))
"
