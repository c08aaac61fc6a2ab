#!/bin/bash
# VAE visit (1 GPU): unit + parity tests of the VAE kernels, then full-size timings.
TAG=${1:-r02v}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_vae_gpu.py -q -s --tb=short --durations=5 > $OUT/${TAG}_pytest_vae.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_vae.log
grep -E "passed|failed|error|rel_l2|ours|Error|assert" $OUT/${TAG}_pytest_vae.log | head -60
timeout 700 python tools/bench_vae.py --out $OUT/${TAG}_vae_bench.json ${2:-} > $OUT/${TAG}_vae_bench.log 2>&1; echo "bench rc=$?" >> $OUT/${TAG}_vae_bench.log
tail -25 $OUT/${TAG}_vae_bench.log
