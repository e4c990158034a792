"""Data parallelism: one process per GPU, the batch sharded across ranks, and a
single all-reduce per step over the flat gradient buffer (torch.distributed,
NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference is single-process; its semantics per
replica are kept: BatchNorm moments stay local to a replica's shard, while the
TALR moments (mean p_tr^2, mean p_tr per tree node, lib/net_types.py:25-27)
ride in the tail of the all-reduced buffer so every replica applies the same
learning-rate scales and the parameters stay bit-identical across ranks.
"""
import os

import torch
import torch.distributed as dist

__all__ = ['init_from_env', 'shard', 'allreduce_flat_', 'experiment_shards', 'run_experiment_parallel']


def init_from_env(backend=None, device=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).
    Returns (rank, world).  A world of 1 needs no process group."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {'device_id': device} if (backend == 'nccl' and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def shard(x, rank, world):
    """Contiguous shard `rank` of the leading (example) axis; the batch must divide evenly
    so every replica sees the reference's per-replica batch size."""
    n = len(x)
    if n % world:
        raise ValueError('batch %d does not divide over %d ranks' % (n, world))
    per = n // world
    return x[rank * per:(rank + 1) * per]


def allreduce_flat_(flat):
    """Sum `flat` ([gradients | per-node TALR moments]) over all ranks in place.  The 1/world
    factor is applied by the optimiser kernel (hyp[GSCALE]), not here."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


# --------------------------------------------------------------------------- #
# Experiment parallelism: an experiment is a list of independent nets -- eight k_cpt
# values or eight depths (scripts/train-nets:29-88) -- that the reference trains one after the other
# (train-nets:159-164).  One net per GPU needs no communication at all.
# --------------------------------------------------------------------------- #
def experiment_shards(net_indices, devices):
    """round-robin assignment {device: [net index, ...]}; devices without work are dropped"""
    net_indices, devices = list(net_indices), list(devices)
    if not devices:
        raise ValueError('no devices')
    out = {d: net_indices[k::len(devices)] for k, d in enumerate(devices)}
    return {d: v for d, v in out.items() if v}


def run_experiment_parallel(argv, net_indices, devices, python=None):
    """Re-run the calling driver once per device with `--nets <its share>` and CUDA_VISIBLE_DEVICES set;
    returns the largest exit code.  `argv` is the driver's own command line without --devices / --nets."""
    import subprocess
    import sys
    procs = []
    for d, share in experiment_shards(net_indices, devices).items():
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(d))
        cmd = [python or sys.executable] + list(argv) + ['--nets'] + [str(i) for i in share]
        procs.append((d, subprocess.Popen(cmd, env=env)))
    return max((p.wait() for _, p in procs), default=0)
