#!/bin/bash
# Round-2 first GPU visit: attention A/B, full parity suite, bench (both arms), same-GPU eager-oracle context.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/r02a_gpu.txt 2>&1
timeout 600 python tools/gpu_check.py attn_ab > $OUT/r02a_attn_ab.log 2>&1; echo "attn_ab rc=$?" >> $OUT/r02a_attn_ab.log
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=15 > $OUT/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/r02a_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --breakdown > $OUT/r02a_bench.json 2> $OUT/r02a_bench.err; echo "bench rc=$?" >> $OUT/r02a_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02a_bench_reference.json 2>> $OUT/r02a_bench.err
timeout 600 python tools/gpu_oracle_baseline.py > $OUT/r02a_gpu_eager_oracle.json 2> $OUT/r02a_gpu_eager_oracle.err
tail -25 $OUT/r02a_attn_ab.log; tail -30 $OUT/r02a_pytest_gpu.log; tail -3 $OUT/r02a_bench.err; cat $OUT/r02a_gpu_eager_oracle.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02a_bench.json')); print({k:d[k] for k in ('value','ms_per_step','achieved_tflops_per_gpu','frac_of_dense_bf16_spec_2250','clocks')}); print(d['e2e']); print(d['roofline']); print(d.get('cpu_baseline'))
PY
