# One GPU call: A/B of compile-time variants (dev tool; results of the last run: profiles/r2m_ab_probes.md).
#   here (no GPU):  bash tools/ab_r2.sh build         -> probe binaries + variant libraries under tools/_build/
#   on the GPU box: bash tools/ab_r2.sh run > gpurun_out/ab.log
# Every binary runs under its own `timeout`: a variant that hangs (e.g. a setmaxnreg.inc beyond the CTA's register pool) costs
# seconds, not the call.
I="-I pytorch-distributed-reinforcement-learning_b200/csrc -I include"
N="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 $I"
if [ "$1" = build ]; then
  mkdir -p tools/_build
  $N -DB200RL_BWD_NI=2 -DB200RL_BWD_NI_TORSO=2 tools/bwd_prof.cu -o tools/_build/bwd_time_ni2 &
  $N -DB200RL_BWD_NI=4 -DB200RL_BWD_NI_TORSO=4 tools/bwd_prof.cu -o tools/_build/bwd_time_ni4 &
  $N -DB200RL_BWDPROF tools/bwd_prof.cu -o tools/_build/bwd_prof &
  $N -DB200RL_FWD_HEADS_PROD=0 tools/fwd3_prof.cu -o tools/_build/fwd3_time_hp0 &
  $N -DB200RL_FWD_HEADS_PROD=1 tools/fwd3_prof.cu -o tools/_build/fwd3_time_hp1 &
  $N -DB200RL_FWD3PROF tools/fwd3_prof.cu -o tools/_build/fwd3_prof &
  for m in 0 1 2; do $N -DB200RL_WGPROBE=$m tools/wgrad_prof.cu -o tools/_build/wgrad_p$m & done
  $N -DB200RL_WG_BF16X=1 tools/wgrad_prof.cu -o tools/_build/wgrad_x1 &
  wait
  exit 0
fi
cd tools/_build
for x in bwd_time_ni2 bwd_time_ni4; do for m in 2 0; do timeout 15 ./$x $m || echo "$x $m: rc $?"; done; done
timeout 15 ./bwd_prof 2
for x in fwd3_time_hp0 fwd3_time_hp1; do timeout 15 ./$x || echo "$x: rc $?"; timeout 15 ./$x 0; done
timeout 15 ./fwd3_prof
for x in wgrad_p0 wgrad_x1 wgrad_p1 wgrad_p2; do timeout 15 ./$x || echo "$x: rc $?"; done
cd ../..
for t in 1 0; do echo "== B200RL_TILE16=$t, B=2048"; B200RL_TILE16=$t timeout 90 python tools/kernel_times.py 2048 128 2>&1 | tail -2 | head -1; done
echo "== default library, B=4096"; timeout 90 python tools/kernel_times.py 2>&1 | tail -2 | head -1
