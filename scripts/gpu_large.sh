mkdir -p gpurun_out
timeout 900 python bench.py --workload synth255 --walkers 8192 --steps 5 --warmup 3 > gpurun_out/bench_synth255.log 2> gpurun_out/bench_synth255.err; echo "rc=$?" >> gpurun_out/bench_synth255.err
timeout 900 python bench.py --workload synth511 --walkers 8192 --steps 3 --warmup 3 > gpurun_out/bench_synth511.log 2> gpurun_out/bench_synth511.err; echo "rc=$?" >> gpurun_out/bench_synth511.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
