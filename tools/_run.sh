timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_shared_mesh_nccl.py > gpurun_out/r2k_check.txt 2>&1
echo "rc=$?" >> gpurun_out/r2k_check.txt
tail -12 gpurun_out/r2k_check.txt
for f in 1 0; do
NR_FUSED_ALLREDUCE=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload cfg3 --steps 300 --warmup 10 > gpurun_out/r2k_cfg3_n2_fused$f.json 2> gpurun_out/r2k_cfg3_n2_fused$f.err
tail -c 400 gpurun_out/r2k_cfg3_n2_fused$f.json
done
