#!/bin/bash
# Round-2 evidence refresh after the CTA-pair attention kernel: full GPU suite, bench, ncu launch list + full block capture.
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -s --durations=5 > $OUT/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/r02d_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02d_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/r02d_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --breakdown > $OUT/r02d_bench.json 2> $OUT/r02d_bench.err; echo "bench rc=$?" >> $OUT/r02d_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02d_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r02d_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gf_ -c 40 -f -o $OUT/r02d_block \
    python bench.py --steps 1 --warmup 0 --layers 1 --controlnet-layers 0 --no-e2e --no-cpu-baseline > $OUT/r02d_ncu_full.log 2>&1
grep -v "Warn\|warn\|^$" $OUT/r02d_pytest_gpu.log | tail -12; tail -2 $OUT/r02d_smoke.log; tail -2 $OUT/r02d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02d_bench.json')); print({k:d[k] for k in ('value','ms_per_step','achieved_tflops_per_gpu','frac_of_dense_bf16_spec_2250','clocks')}); print(d['e2e']); print(d['roofline']['achieved'], d['roofline']['frac'])
PY
