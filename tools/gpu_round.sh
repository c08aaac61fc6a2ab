#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list, ncu full capture of one block.
# Usage (from the authoring box): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/${TAG}_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 --breakdown > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?" >> $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gf_ -c 40 -f -o $OUT/${TAG}_block \
    python bench.py --steps 1 --warmup 0 --layers 1 --controlnet-layers 0 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
timeout 600 python tools/bench_vae.py --out $OUT/${TAG}_vae_bench.json > $OUT/${TAG}_vae_bench.log 2>&1
tail -5 $OUT/${TAG}_pytest_gpu.log; cat $OUT/${TAG}_smoke.log | tail -3; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
