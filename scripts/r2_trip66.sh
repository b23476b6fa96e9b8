#!/bin/bash
# 2 GPUs: frame-sharded MU (ED + KL, MIN all-reduce) against the single-GPU solve; MU strong scaling point; the 2-GPU test
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_snmf_check.py 2>&1 | grep dist_snmf | tee gpurun_out/t66_dist.txt
MU_ITERS=20 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 scripts/mu_scaling.py 2>/dev/null | cut -c1-300 | tee -a gpurun_out/t66_dist.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "frame_sharded" 2>&1 | tail -2
