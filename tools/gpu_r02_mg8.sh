#!/bin/bash
# Round-2 8-GPU visit: bit-identity tests at world 8, configs[2]/[3]/[4] runs, scaling point at N = 8.
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -k "8 or long" > $OUT/r02_mg8_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_mg8_pytest.log
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/r02_bench_8gpu_peer.json 2> $OUT/r02_bench_8gpu_peer.err
timeout 400 $TR tools/bench_configs.py --config sample --steps 40 --cfg-parallel > $OUT/r02_cfg3_cfg2sp4.json 2> $OUT/r02_cfg3_cfg2sp4.err
timeout 600 $TR tools/bench_configs.py --config direct --steps 40 --cfg-parallel --replicas 1 > $OUT/r02_cfg4_direct_cfg2sp4.json 2> $OUT/r02_cfg4_direct_cfg2sp4.err
timeout 400 $TR tools/bench_configs.py --config direct --steps 10 --cfg-parallel --replicas 4 > $OUT/r02_cfg4_direct_4xcfg2.json 2> $OUT/r02_cfg4_direct_4xcfg2.err
timeout 400 $TR tools/bench_configs.py --config long --frames 121 --steps 3 > $OUT/r02_cfg5_121.json 2> $OUT/r02_cfg5_121.err
tail -8 $OUT/r02_mg8_pytest.log; for f in r02_bench_8gpu_peer r02_cfg3_cfg2sp4 r02_cfg4_direct_cfg2sp4 r02_cfg4_direct_4xcfg2 r02_cfg5_121; do echo "== $f"; cut -c1-600 $OUT/$f.json; tail -2 $OUT/$f.err | cut -c1-300; done
