timeout 600 python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r1_bench_reference_arm.json 2>> gpurun_out/r1_bench_n1.err
timeout 600 python bench.py --flush-l2 --e2e-threads 1 --no-cpu > gpurun_out/r1_bench_n1_flush_1thread.json 2>> gpurun_out/r1_bench_n1.err
timeout 600 python tools/engine_bench.py --seconds 5 --moves 4 --threads 16 > gpurun_out/r1_engine_bench.json 2>&1
timeout 600 python tools/engine_bench.py --netbench --threads 64 >> gpurun_out/r1_engine_bench.json 2>&1
timeout 600 python tools/batch_sweep.py > gpurun_out/batch_sweep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
nproc
