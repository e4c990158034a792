"""Checkpoint / resume — beyond the reference, which only writes the final net.

A checkpoint is one `.npy` holding a dict:
    'net'       exactly what `serdes.write_net` writes (so `read_net`-style consumers and the
                `nets/<expt>/NNNN.npy` layout of scripts/train-nets:149-157 keep working on it),
    'momentum'  the optimiser's accumulators, one array per trainable tensor in the engine's
                enumeration order (the reference's MomentumOptimizer slots, net_types.py:36),
    'step'      number of completed training steps (drives the lambda / tau schedules),
    'rng'       state of the data-sampling numpy Generator, if given.
Running statistics of BatchNorm are ordinary (non-trainable) parameters and travel inside 'net'.
"""
import numpy as np

from lib.serdes import decode_net, encode_net

__all__ = ['save_checkpoint', 'load_checkpoint', 'net_record']

FORMAT = 1


def save_checkpoint(path, net, step, rng=None):
    eng = getattr(net, '_engine', None)
    np.save(path, {
        'format': FORMAT,
        'net': encode_net(net),
        'momentum': eng.momentum_numpy() if eng is not None else None,
        'step': int(step),
        'rng': rng.bit_generator.state if rng is not None else None})


def load_checkpoint(path, rng=None, **configure):
    """-> (net, step).  `configure` is forwarded to net.configure (precision=..., graphs=...);
    the momentum is restored as soon as the net owns an engine (needs the GPU)."""
    d = np.load(path, allow_pickle=True)[()]
    if d.get('format') != FORMAT:
        raise ValueError('%s: not a checkpoint (format %r)' % (path, d.get('format')))
    net = decode_net(d['net'])
    if configure:
        net.configure(**configure)
    if d['momentum'] is not None:
        net._pending_momentum = d['momentum']       # consumed by Net._get_engine
    if rng is not None and d['rng'] is not None:
        rng.bit_generator.state = d['rng']
    return net, int(d['step'])


def net_record(path):
    """the `write_net` payload of a checkpoint (or of a plain write_net file)"""
    d = np.load(path, allow_pickle=True)[()]
    return d['net'] if 'format' in d else d
