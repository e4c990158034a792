"""Latency of the data-parallel step tail in isolation (run under torchrun, one rank per GPU):
ncclAllReduce (whole buffer / last bucket) + optimiser against the fused peer-memory kernel (whole vector /
shallow slice / deep slice).  Tuning aid, not a bench line."""
import ctypes, datetime, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')]
import numpy as np, torch
import torch.distributed as dist
import bench

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', rank)))
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=120))
B = int(os.environ.get('B', 128))


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


out = []
for fused in (0, 1):
    os.environ['MPNN_DIST_FUSED'] = str(fused)
    run = bench.Run('cifar10-ac', B, 'bf16', dev, rank, world, graphs=False)
    eng, plan = run.eng, run.plan
    run.warm(3)
    eng.stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    split = plan.ar_split
    if not fused:
        out.append(('nccl allreduce, whole buffer (%d floats)' % eng.grad.numel(), timeit(lambda: eng._allreduce(0, None))))
        out.append(('nccl allreduce, last bucket (%d floats)' % split, timeit(lambda: eng._allreduce(0, split))))
        out.append(('optimiser kernel', timeit(lambda: plan.opt_ops[0]())))
        out.append(('nccl last bucket + optimiser', timeit(lambda: (eng._allreduce(0, split), plan.opt_ops[0]()))))
    else:
        out.append(('fused p2p, whole vector', timeit(lambda: plan._p2p_tail(0, 0, 0))))
        out.append(('fused p2p, shallow slice + moments (channel 0)', timeit(lambda: plan._p2p_tail(0, split - eng.g0, 0))))
        out.append(('fused p2p, deep slice (channel 1)', timeit(lambda: plan._p2p_tail(split - eng.g0, 0, 1))))
        assert eng.p2p_status() == 0
if rank == 0:
    print('# %d ranks, cifar10-ac, %d parameters; us per call, max over ranks, back to back on one stream' % (world, eng.n_theta))
    for k, v in out:
        print('%-60s %8.1f us' % (k, v))
dist.barrier()
dist.destroy_process_group()
