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
import os

import numpy as np

from lib.serdes import decode_net, encode_net

__all__ = ['save_checkpoint', 'load_checkpoint', 'net_record']

FORMAT = 1


def save_checkpoint(path, net, step, rng=None, keep_previous=True):
    """Atomic: the payload is written to a temporary file in the same directory and moved onto `path`
    with os.replace, so a crash mid-write never destroys the checkpoint `--resume` reads; the previous
    checkpoint is kept beside it as `<path>.prev` unless keep_previous is False."""
    eng = getattr(net, '_engine', None)
    payload = {
        'format': FORMAT,
        'net': encode_net(net),
        'momentum': eng.momentum_numpy() if eng is not None else None,
        'step': int(step),
        'rng': rng.bit_generator.state if rng is not None else None}
    tmp = '%s.tmp.%d' % (path, os.getpid())
    with open(tmp, 'wb') as f:
        np.save(f, payload)
        f.flush()
        os.fsync(f.fileno())
    if keep_previous and os.path.exists(path):
        os.replace(path, path + '.prev')
    os.replace(tmp, path)


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


def _walk(rec, path=''):
    """(path, layer record) for every layer record of a net record, preorder"""
    yield path, rec
    for i, c in enumerate(rec.get('comps', [])):
        yield from _walk(c, '%s.comps[%d]' % (path, i))
    if rec.get('router') is not None:
        yield from _walk(rec['router'], path + '.router')
    for i, s in enumerate(rec.get('sinks', [])):
        yield from _walk(s, '%s/%d' % (path, i))


def describe(path):
    """text summary of a write_net file or checkpoint: net type, hypers, every parameter tensor"""
    d = np.load(path, allow_pickle=True)[()]
    rec = d['net'] if 'format' in d else d
    lines = ['%s: %s' % (path, rec['type'])]
    if 'format' in d:
        lines.append('  checkpoint: step %d, momentum %s, sampler rng %s' % (
            d['step'], 'yes' if d['momentum'] is not None else 'no', 'yes' if d['rng'] is not None else 'no'))
    lines.append('  hypers: ' + ', '.join('%s=%r' % kv for kv in sorted(rec['hypers'].items())))
    n = 0
    for p, layer in _walk(rec['root'], 'root'):
        for k, v in sorted(layer.get('params', {}).items()):
            v = np.asarray(v)
            n += v.size
            lines.append('  %-44s %-10s %-18s |x| %.4g' % (p + ':' + layer.get('type', '?'), k, tuple(v.shape), float(np.linalg.norm(v))))
    lines.append('  %d parameters (running BatchNorm moments included)' % n)
    return '\n'.join(lines)


def roundtrip(path):
    """decode -> encode a net file and compare every array bit for bit (read_net / write_net consistency)"""
    d = np.load(path, allow_pickle=True)[()]
    rec = d['net'] if 'format' in d else d
    again = encode_net(decode_net(rec))
    assert again['type'] == rec['type'] and again['hypers'] == dict(rec['hypers'])
    a, b = list(_walk(rec['root'], 'root')), list(_walk(again['root'], 'root'))
    assert [p for p, _ in a] == [p for p, _ in b], 'topology changed'
    for (p, la), (_, lb) in zip(a, b):
        assert sorted(la.get('params', {})) == sorted(lb.get('params', {})), p
        for k in la.get('params', {}):
            if not np.array_equal(np.asarray(la['params'][k]), np.asarray(lb['params'][k])):
                raise AssertionError('%s:%s differs after the round trip' % (p, k))
    return len(a)


if __name__ == '__main__':
    import sys
    if len(sys.argv) != 3 or sys.argv[1] not in ('info', 'roundtrip'):
        raise SystemExit('usage: python -m lib.checkpoint info|roundtrip FILE.npy')
    if sys.argv[1] == 'info':
        print(describe(sys.argv[2]))
    else:
        print('%s: %d layer records identical after decode -> encode' % (sys.argv[2], roundtrip(sys.argv[2])))
