# One GPU call (round 2): ncu launch list of the bench command, one --set full capture of every kernel of a PPO update, one of
# the scan / V-MPO / SAC kernels, CSV exports.  Usage on the GPU box:  bash tools/gpu_prof_r2.sh r2g
set -x
TAG=${1:-r2x}
mkdir -p gpurun_out
K='regex:lstm_|head_|ppo_head|impala_head|peer_step|reduce_partials|torso_grad'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s ${SKIP:-18} -c ${CNT:-36} --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s ${SKIP:-18} -c ${PER:-6} -o gpurun_out/prof_${TAG} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}.csv 2>/dev/null
SCAN_BENCH_ITERS=1 timeout 900 ncu --set full --clock-control none -k 'regex:gae_kernel|td_gae|vtrace|vmpo_|sac_|clip_adam|soft_update|kl_kernel' -c 60 -o gpurun_out/prof_scans_${TAG} \
    python tools/scan_bench.py > gpurun_out/ncu_scans_${TAG}.log 2>&1
ncu -i gpurun_out/prof_scans_${TAG}.ncu-rep --page raw --csv > gpurun_out/raw_scans_${TAG}.csv 2>/dev/null
timeout 300 python tools/scan_bench.py > gpurun_out/scan_bench_${TAG}.jsonl 2> gpurun_out/scan_bench_${TAG}.err
rm -f gpurun_out/prof_scans_${TAG}.ncu-rep
ls -la gpurun_out | tail -12
