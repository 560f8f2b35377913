mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k3_szmap|k2_dgemm|k1_profiles" -s 6 -c 4 -f -o gpurun_out/final_full python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_final.log 2>&1
