set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python bench.py 2>&1 | tee gpurun_out/bench_r1a.json | tail -3
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -2
# full ncu capture of the three LSTM kernels (one launch each, after 2 warm-up updates)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_ -s 6 -c 3 -o gpurun_out/prof_r1a python tools/kernel_times.py 4096 128 4 2 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
