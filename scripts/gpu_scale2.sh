mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
