# One GPU call: parity tests, the bench line, the ncu launch list of the bench command, one --set full capture of the
# main kernels.  Usage on the GPU box:  bash tools/gpu_prof.sh r1h
set -x
TAG=${1:-r1x}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py 2> gpurun_out/bench_${TAG}.err | tee gpurun_out/bench_${TAG}.json | tail -1 | cut -c1-600
timeout 300 python tools/bench_configs.py > gpurun_out/configs_${TAG}.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 66 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lstm_|torso_grad|head_grad|head_fwd' -s 21 -c 14 -o gpurun_out/prof_${TAG} \
    python tools/kernel_times.py 4096 128 4 2 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}.csv 2>/dev/null
ls -la gpurun_out | tail -12
