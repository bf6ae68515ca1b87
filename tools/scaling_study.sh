#!/bin/bash
# Scaling study on one 8-GPU box (gpurun --gpus 8 -- 'bash tools/scaling_study.sh'): strong-scaled config 2,
# weak-scaled config 3 with the fused exchange and with NCCL, and the shared-mesh gradient check on 8 ranks.
# The lines it appends to gpurun_out/r3u_*.jsonl are what profiles/r2_scaling_*.jsonl hold.
run() { # n workload extra-env out
  env $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $1 --workload $2 --steps 300 --warmup 10 2>> gpurun_out/r3u_multi.err | tail -1 >> $4
}
python tools/check_shared_mesh_nccl.py > /dev/null 2>&1 || true
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29477 tools/check_shared_mesh_nccl.py > gpurun_out/r3u_check8.txt 2>&1
for n in 8 4 2; do
  run $n cfg2 "NR_X=0" gpurun_out/r3u_cfg2_strong.jsonl
  run $n cfg3 "NR_FUSED_ALLREDUCE=1" gpurun_out/r3u_cfg3_fused.jsonl
  run $n cfg3 "NR_FUSED_ALLREDUCE=0" gpurun_out/r3u_cfg3_nccl.jsonl
done
python bench.py --workload cfg2 --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r3u_cfg2_strong.jsonl
python bench.py --workload cfg3 --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/r3u_cfg3_fused.jsonl
nvidia-smi topo -m > gpurun_out/r3u_topo.txt 2>&1
