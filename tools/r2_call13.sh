#!/bin/bash
# Round 2, call 13: FiLM chunk size sweep (GDL_FILM_CHUNK_I) with the unrolled contraction.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -k "film" > gpurun_out/r2c13_tests.log 2>&1
echo "== film tests exit $?"; grep -E "passed|failed|^FAILED" gpurun_out/r2c13_tests.log | tail -4
for ch in 64 128 256 512; do
  GDL_FILM_CHUNK_I=$ch timeout 300 python bench.py --fusion film --steps 10 --warmup 3 --no-cpu --no-device-pipeline > gpurun_out/r2c13_bench_film$ch.log 2>&1
  echo "== bench film chunk $ch exit $?"; grep '^{"metric"' gpurun_out/r2c13_bench_film$ch.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); kb = d['kernel_breakdown']
print('   ms/step %.3f value %.0f launches %d' % (d['ms_per_step'], d['value'], d['launches_per_step']))
print('   ' + ' '.join('%s=%.2f' % (k, v['ms']) for k, v in kb.items() if k in ('film_outer','film_contract','transpose','gemm_nt','gemm_tn','sgd_momentum','grad_stats')))
"
done
