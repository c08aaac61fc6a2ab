#!/bin/bash
# Short GPU visit: parity tests + bench (no ncu).  gpurun --timeout 1500 -- 'bash tools/gpu_quick.sh <tag> [bench args]'
TAG=${1:-quick}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --breakdown "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" >> $OUT/${TAG}_bench.err
tail -4 $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_bench.err | tail -12; python -c "
import json,sys
d=json.load(open('$OUT/${TAG}_bench.json')); print({k:d[k] for k in ('value','ms_per_step','achieved_tflops_per_gpu','frac_of_dense_bf16_spec_2250','clocks')}); print(d['e2e'])"
