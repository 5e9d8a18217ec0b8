"""Where the data-parallel step time goes (dev tool; launch with torch.distributed.run, one rank per GPU):
each rank alone (no exchange), the fused peer step alone in lock-step, and the full data-parallel update."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def timeit(fn, n, dev):
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / n


def main(B=4096, T=128, iters=100):
    from b200rl.engine import OnPolicyEngine
    from b200rl.networks import models as M
    from b200rl.synth import make_rollouts
    from tools.kernel_times import HYPER
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    bs = [{k: v.to(dev) for k, v in make_rollouts(B, T, 4, 2, batch_index=rank * 10 + i).items()} for i in range(4)]
    res = {"rank": rank}
    # 0. before any process group / peer mapping exists
    solo = OnPolicyEngine(M.MlpLSTMSingle(4, 2, T, 64).to(dev).actor, "PPO", HYPER)
    for i in range(5):
        solo.update(bs[i % 4])
    res["before_init_ms"] = timeit(lambda i: solo.update(bs[i % 4]), iters, dev)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    res["after_nccl_init_ms"] = timeit(lambda i: solo.update(bs[i % 4]), iters, dev)
    # 1. this GPU alone
    dist.barrier()
    for r in range(world):                      # one GPU at a time, the others idle
        if r == rank:
            res["alone_ms"] = timeit(lambda i: solo.update(bs[i % 4]), iters, dev)
        dist.barrier()
    res["solo_ms"] = timeit(lambda i: solo.update(bs[i % 4]), iters, dev)      # all GPUs at once, no exchange
    # 2. data parallel
    eng = OnPolicyEngine(M.MlpLSTMSingle(4, 2, T, 64).to(dev).actor, "PPO", HYPER, process_group=dist.group.WORLD)
    for i in range(5):
        eng.update(bs[i % 4])
    dist.barrier()
    res["dp_ms"] = timeit(lambda i: eng.update(bs[i % 4]), iters, dev)
    # 3. the peer step alone, ranks in lock-step
    ch = eng.chan
    st = eng.ops._stream(dev)
    sp, ss, sg = eng.flat.clone(), eng.sq_avg.clone(), torch.empty_like(eng.grad)

    def step(i):
        ch.begin()
        ch.step(0, sp, ss, None, None, sg, eng.scalars, eng.grad_norm, 1.0 / world, 40.0, 1e-4, 0.99, 0.0, 1e-5, 0, st)
    for i in range(5):
        step(i)
    dist.barrier()
    res["peer_step_ms"] = timeit(step, 200, dev)
    ch.check()
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        print(json.dumps({"world": world, "ranks": out}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(*[int(x) for x in sys.argv[1:]])
