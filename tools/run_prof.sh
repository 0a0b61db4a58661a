python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/trace_items.py both 256 > gpurun_out/r2a_trace_items.txt 2>&1; tail -40 gpurun_out/r2a_trace_items.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/r2a_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trunk_kernel -s 4 -c 1 -o gpurun_out/trunk_r2a -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2a_ncu_full.log 2>&1
ls -la gpurun_out/trunk_r2a.ncu-rep
