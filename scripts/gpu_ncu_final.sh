mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k3_szmap|k7_filter|k2_dgemm|k1_profiles" -s 8 -c 5 -f -o gpurun_out/final_full python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_final.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
nvidia-smi > gpurun_out/nvidia_smi.txt; lscpu | head -20 > gpurun_out/lscpu.txt
