python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/sweep_opts.py 2>&1 | tail -9
ncu --set full --clock-control none --import-source on -k regex:trunk_kernel -s 4 -c 1 -o gpurun_out/trunk_r2b -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2b_ncu_full.log 2>&1
ls -la gpurun_out/trunk_r2b.ncu-rep
