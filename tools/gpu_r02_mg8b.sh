#!/bin/bash
# 8-GPU visit 2: ControlNet branch on a second stream (bit identity at world 2 and 8, bench on / off).
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s -k "peer-2 or peer-8" > $OUT/r02_mg8b_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02_mg8b_pytest.log
for m in off on off on; do
  timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 --controlnet-stream $m > $OUT/r02_bench_8gpu_cn_$m.json 2> $OUT/r02_bench_8gpu_cn_$m.err
  python -c "
import json
d=json.load(open('$OUT/r02_bench_8gpu_cn_$m.json')); print('cn-stream $m', d['ms_per_step'], d['clocks']['sm_mhz'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if k in ('peer_barrier','rmsnorm_rope','attention_self','gemm')})"
done
grep -v "Warn\|warn\|fork\|^$\|Docs" $OUT/r02_mg8b_pytest.log | tail -6
