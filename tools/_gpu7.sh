set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_v1.json 2> gpurun_out/r2_bench_n2_v1.err
echo rc=$?
tail -c 1500 gpurun_out/r2_bench_n2_v1.err
head -c 2500 gpurun_out/r2_bench_n2_v1.json
