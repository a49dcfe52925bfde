#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; ( "$@" ) > gpurun_out/$name.log 2>&1; echo "== $name exit $?"; tail -15 gpurun_out/$name.log; }
run t_film_k env timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm or film" --no-header -p no:cacheprovider -x
run t_film_s env timeout 900 python -m pytest tests/test_gpu_step.py -q -m gpu -k "film" --no-header -p no:cacheprovider -x
run t_film_c env timeout 900 python -m pytest tests/test_gpu_compat.py -q -m gpu --no-header -p no:cacheprovider -x
run b_film env GDL_DUMP_OPS=gpurun_out/ops_film.json timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --fusion film
