# One GPU call: A/B of compile-time variants (dev tool).  bash tools/ab_r2.sh > gpurun_out/ab.log
cd tools/_build
for x in bwd_time_ni2 bwd_time_ni4; do for m in 0 2; do timeout 60 ./$x $m; done; done
timeout 60 ./bwd_prof_ni2 2; timeout 60 ./bwd_prof_ni4 2
timeout 60 ./fwd3_time; timeout 60 ./fwd3_prof
for x in wgrad_x0 wgrad_x1 wgrad_p1 wgrad_p2; do timeout 60 ./$x; done
cd ../..
for tag in default ni4 x1 ni4x; do
  L=tools/_build/lib_$tag/libb200rl.so; [ $tag = default ] && L=
  echo "== $tag"; B200RL_LIB=$L timeout 120 python tools/kernel_times.py 2>&1 | tail -2
done
echo "== parity with ni4x"
B200RL_LIB=tools/_build/lib_ni4x/libb200rl.so timeout 400 python -m pytest tests/test_baseline_shapes.py tests/test_learner_step.py -m gpu -x -q 2>&1 | tail -5
