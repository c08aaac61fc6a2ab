#!/bin/bash
# Final evidence visit of the round (1 GPU): full GPU suite, smoke, DiT bench (+ reference arm), VAE bench, ncu captures
# of the convolution kernels, launch list of a VAE decode + encode.
TAG=${1:-r02z}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --breakdown > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" >> $OUT/${TAG}_bench.err
timeout 600 python tools/bench_vae.py --out $OUT/${TAG}_vae_bench.json > $OUT/${TAG}_vae_bench.log 2>&1; echo "vae bench rc=$?" >> $OUT/${TAG}_vae_bench.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gf_conv3d -c 4 -f -o $OUT/${TAG}_conv python tools/conv_bench.py --shapes tile_s3,s2,s1,enc1 --iters 1 > $OUT/${TAG}_conv_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gf_conv3d -c 2 -f -o $OUT/${TAG}_conv_fused python tools/conv_bench.py --epi resnorm --shapes tile_s3,s2 --iters 1 >> $OUT/${TAG}_conv_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_vae_launches.csv python tools/bench_vae.py --launch-list > $OUT/${TAG}_vae_launches.log 2>&1
tail -4 $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_bench.err; grep -v "^{" $OUT/${TAG}_vae_bench.log | tail -10
python - <<PY
import json
d=json.load(open('$OUT/${TAG}_bench.json')); print({k:d[k] for k in ('value','ms_per_step','achieved_tflops_per_gpu','frac_of_dense_bf16_spec_2250','clocks')}); print(d['e2e']); print(d['roofline']['achieved'], d['roofline']['frac'])
PY
