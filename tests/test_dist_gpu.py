"""Two-rank data parallelism on real GPUs (NCCL): the product path -- `net.configure(dist=True)`,
`lib.parallel.shard`, one all-reduce per step over [gradients | TALR moments] -- against the oracle.

  * after 3 steps `theta`, the momentum accumulators and the reduced buffer are BIT-identical across ranks --
    with the NCCL all-reduce followed by the optimiser, and with the fused tail (MPNN_DIST_FUSED=1: ONE kernel that
    reduce-scatters the gradients by peer loads, applies TALR + momentum to its slice and all-gathers the new
    parameters by peer stores over NVLink, csrc/p2p.cu);
  * after the first step the parameters equal the oracle's data-parallel step on the two shards
    (per-replica BatchNorm, averaged gradients, TALR moments averaged over ranks; the restatement of
    csrc/optim.cu in tests/test_dist_cpu.py).

Needs two GPUs (skipped otherwise): run with `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

from test_dist_cpu import _free_port, _local, _update  # noqa: E402
from util import batch, randomize_routers, record_of, tiny_net  # noqa: E402

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, rec, x0, y, q, graphs, fused=False):
    os.environ['MPNN_DIST_FUSED'] = '1' if fused else '0'
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, 'multipath-nn_b200'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from lib import parallel, serdes
    torch.cuda.set_device(rank)
    r, w = parallel.init_from_env('nccl', device=torch.device('cuda', rank))
    net = serdes.decode_net(rec).configure(precision='fp32', dist=True, graphs=graphs)
    xs, ys = parallel.shard(x0, rank, world), parallel.shard(y, rank, world)
    eng = net._get_engine()
    snaps = []
    for t in range(3):
        net.train.run({net.x0: xs, net.y: ys, net.mode: 'tr', net.τ: 0.8, net.λ_lrn: 0.1})
        torch.cuda.synchronize()
        snaps.append((eng.theta.cpu().numpy().copy(), eng.accum.cpu().numpy().copy(), eng.grad.cpu().numpy().copy()))
    assert eng.fused_dp == bool(fused)
    if fused:
        assert eng.p2p_status() == 0, 'a wait inside the fused tail timed out'
        snaps = [(th, ac, g[eng.g0:]) for th, ac, g in snaps]      # the per-node moments are reduced in the kernel, not in place
    q.put((rank, snaps, [(p._bind[2], p.value.size) for p in eng.tparams]))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('fused', [False, True])         # ncclAllReduce + optimiser | ONE kernel over NVLink peer memory (csrc/p2p.cu)
@pytest.mark.parametrize('graphs', [False, True])        # eager launches | the whole step (all-reduce included) as one CUDA graph
def test_two_rank_nccl_replicas_are_bit_identical_and_match_the_oracle(graphs, fused):
    net = randomize_routers(tiny_net('ac', k_cpt=4e-9))
    rec = record_of(net)
    x0, y = batch(32, seed=4)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rec, x0, y, q, graphs, fused)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    (_, s0, layout), (_, s1, _) = res
    for t in range(3):
        for a, b, what in zip(s0[t], s1[t], ('theta', 'momentum', 'reduced [grad | TALR moments]')):
            assert np.array_equal(a, b), ('step %d: %s differs between the ranks' % (t, what))
    # first step against the oracle's data-parallel update (momentum starts at zero)
    parts = [_local(rec, x0[i * 16:(i + 1) * 16], y[i * 16:(i + 1) * 16]) for i in range(2)]
    flat = parts[0][2] + parts[1][2]
    want = _update(parts[0][0], parts[0][1], flat, 2, lr=0.1).numpy()
    got = np.concatenate([s0[0][0][off:off + n] for off, n in layout])
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    print('2-rank step vs oracle: rel err %.2e' % err)
    assert err < 1e-5, err
