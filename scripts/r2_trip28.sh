#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu -x ) > gpurun_out/r2t28_tests.log 2>&1; tail -3 gpurun_out/r2t28_tests.log | cut -c1-300
for bn in 128 256; do
echo "== DRNMF_GEMM_BN=$bn"
DRNMF_GEMM_BN=$bn timeout 900 python bench.py --no-cpu-baseline --no-throughput --no-parity --mu-frames 225000 > gpurun_out/r2t28_bench_bn$bn.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t28_bench_bn$bn.json').read().strip())
print('value %.0f ms %.2f stage %s'%(d['value'], d['ms_per_step'], d['roofline']['stage_ms']))
ex=d['config']['extras']
print('train %.2f ms | mu22k %.3f ms | mu225k %.3f ms'%(ex['training_step']['ms'], ex['snmf_mu_ed_n22528']['ms_per_iteration'], ex['snmf_mu_ed_n225000']['ms_per_iteration']))
PY
done
