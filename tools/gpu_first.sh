set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -c "import torch; print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count)"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/kernel_times.py 4096 128 2>&1 | tail -5
timeout 120 python tools/kernel_times.py 64 32 2>&1 | tail -3
timeout 120 python tools/kernel_times.py 8192 5 2>&1 | tail -3
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1a.csv python tools/kernel_times.py 4096 128 4 4 2 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
