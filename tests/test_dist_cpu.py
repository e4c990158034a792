"""World-size-2 data-parallel plumbing on CPU (gloo): batch sharding, the single
flat all-reduce carrying [gradients | TALR moments], and the update rule that
keeps replicas bit-identical.  The arithmetic on each rank is the oracle's (the
CUDA engine cannot run here); what is under test is lib.parallel and the
reduction / scaling convention the optimiser kernel implements
(csrc/optim.cu: g = grad*gscale + 2*l2*mean(p_tr)*theta, scaled by
mult / sqrt(mean p_tr^2), both moments averaged over ranks)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from util import batch, randomize_routers, record_of, tiny_net


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local(rec, x0, y):
    """per-replica quantities: raw gradients (without L2), per-node moments, parameters"""
    from oracle.torch_ref import OracleNet
    o = OracleNet(rec, torch.float64)
    for _, _, _, t in o.trainable:
        t.grad = None
    out = o.forward(x0, y, 'tr', tau=0.8)
    data = out.c_tot - sum(((nd.p_tr.detach() * (nd.c_mod + (nd.router.c_mod if nd.router is not None else 0.0))).mean()
                            for nd in out.nodes.values()))
    data.backward()
    grads = [t.grad.reshape(-1) if t.grad is not None else torch.zeros(t.numel(), dtype=torch.float64)
             for *_, t in o.trainable]
    mom = torch.stack([torch.stack([(out.nodes[p].p_tr.detach() ** 2).mean(), out.nodes[p].p_tr.detach().mean()])
                       for p in out.order]).reshape(-1)
    return o, out, torch.cat(grads + [mom])


def _update(o, out, flat, world, lr=0.1, mu=0.9, k_l2=1e-4):
    """host restatement of talr_momentum_kernel on the reduced buffer"""
    n_nodes = len(out.order)
    mom = flat[-2 * n_nodes:].reshape(n_nodes, 2) / world
    node = {p: i for i, p in enumerate(out.order)}
    off = 0
    new = []
    for path, role, key, t in o.trainable:
        g = flat[off:off + t.numel()].reshape(t.shape) / world
        off += t.numel()
        l2 = k_l2 if key.startswith('w') else 0.0
        g = g + 2 * l2 * mom[node[path], 1] * t.detach()
        g = g / torch.sqrt(mom[node[path], 0])
        new.append(t.detach() - lr * g)            # first step: momentum buffer starts at zero
    return torch.cat([v.reshape(-1) for v in new])


def _worker(rank, world, port, rec, x0, y, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from lib import parallel
    r, w = parallel.init_from_env('gloo')
    assert (r, w) == (rank, world)
    xs, ys = parallel.shard(x0, rank, world), parallel.shard(y, rank, world)
    o, out, flat = _local(rec, xs, ys)
    parallel.allreduce_flat_(flat)
    theta = _update(o, out, flat, world)
    q.put((rank, flat.numpy(), theta.numpy()))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_allreduce_keeps_replicas_identical():
    net = randomize_routers(tiny_net('ac', k_cpt=4e-9))
    rec = record_of(net)
    x0, y = batch(16, seed=4)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rec, x0, y, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # both ranks hold the same reduced buffer and the same updated parameters
    np.testing.assert_array_equal(res[0][1], res[1][1])
    np.testing.assert_array_equal(res[0][2], res[1][2])
    # and it equals the single-process sum of the two shards' contributions
    parts = [_local(rec, x0[i * 8:(i + 1) * 8], y[i * 8:(i + 1) * 8])[2] for i in range(2)]
    np.testing.assert_allclose(res[0][1], (parts[0] + parts[1]).numpy(), rtol=1e-12, atol=1e-15)


def test_shard_rejects_uneven_batches():
    from lib import parallel
    import pytest
    with pytest.raises(ValueError):
        parallel.shard(np.zeros((10, 3)), 0, 4)
    assert len(parallel.shard(np.zeros((12, 3)), 2, 4)) == 3
