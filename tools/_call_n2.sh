mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_sharded_check.py > gpurun_out/n2_sharded_check.log 2> gpurun_out/n2_sharded_check.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --nrays 20000000 --kirchhoff-n 200000 --config5-wave-n 100000 --config5-wave-m 262144 --no-cpu-baseline > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
