#!/bin/bash
# VAE evidence visit (1 GPU): racecheck of the VAE kernels, ncu full captures of both convolution kernels, ncu launch
# list of one decode + encode, final full-size timings next to the eager baseline.
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py --vae-only > $OUT/${TAG}_racecheck_vae.log 2>&1; echo "racecheck rc=$?" >> $OUT/${TAG}_racecheck_vae.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gf_conv3d -c 3 -f -o $OUT/${TAG}_conv python tools/conv_bench.py --shapes tile_s3,s2,up1 --iters 0 > $OUT/${TAG}_conv_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gf_conv3d -c 2 -f -o $OUT/${TAG}_conv_fused python tools/conv_bench.py --fused --shapes tile_s3,s2 --iters 0 >> $OUT/${TAG}_conv_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_vae_launches.csv python tools/bench_vae.py --launch-list > $OUT/${TAG}_vae_launches.log 2>&1
timeout 600 python tools/bench_vae.py --out $OUT/${TAG}_vae_bench.json > $OUT/${TAG}_vae_bench.log 2>&1; echo "bench rc=$?" >> $OUT/${TAG}_vae_bench.log
grep -E "RACECHECK SUMMARY|rc=|ALL OK|FAIL" $OUT/${TAG}_racecheck_vae.log | tail -5
grep -v "^{" $OUT/${TAG}_vae_bench.log | tail -12
