"""Net (de)serialisation in the reference's on-disk format
(/root/reference/scripts/lib/serdes.py:13-60): `np.save` of a
nested dict -- net = {type, root, hypers, params}; layer = {type, name,
hypers, params{key: float32 ndarray}, sinks, comps, router}.  Files written
by either implementation load in the other (NumPy >= 1.16.3 needs
allow_pickle=True, which the reference's `np.load(path)[()]` predates).
"""
import numpy as np

import lib.layer_types
import lib.net_types

__all__ = ['encode_layer', 'decode_layer', 'load_params', 'encode_net', 'decode_net',
           'write_net', 'read_net']


def encode_layer(layer):
    if layer is None:
        return None
    return {
        'type': type(layer).__name__,
        'name': layer.name,
        'hypers': dict(vars(layer.hypers)),
        'params': {k: v.eval() for k, v in vars(layer.params).items()},
        'sinks': [encode_layer(s) for s in layer.sinks],
        'comps': [encode_layer(c) for c in layer.comps],
        'router': encode_layer(layer.router)}


def decode_layer(record):
    if record is None:
        return None
    cls = getattr(lib.layer_types, record['type'])
    return cls(name=record['name'],
               router=decode_layer(record['router']),
               sinks=[decode_layer(r) for r in record['sinks']],
               comps=[decode_layer(r) for r in record['comps']],
               **dict(record['hypers']))


def load_params(layer, record):
    """Assign every stored array to the linked layer tree (positional zip over
    comps / sinks, exactly like serdes.py:27-34)."""
    if layer is None:
        return
    load_params(layer.router, record['router'])
    for sub, rec in zip(layer.comps, record['comps']):
        load_params(sub, rec)
    for sub, rec in zip(layer.sinks, record['sinks']):
        load_params(sub, rec)
    for k, v in record['params'].items():
        getattr(layer.params, k).assign(v)


def encode_net(net):
    return {
        'type': type(net).__name__,
        'root': encode_layer(net.root),
        'hypers': dict(vars(net.hypers)),
        'params': {k: v.eval() for k, v in vars(net.params).items()}}


def decode_net(record):
    cls = getattr(lib.net_types, record['type'])
    net = cls(root=decode_layer(record['root']), **dict(record['hypers']))
    load_params(net.root, record['root'])
    for k, v in record['params'].items():
        getattr(net.params, k).assign(v)
    return net


def write_net(path, net):
    np.save(path, encode_net(net))


def read_net(path):
    return decode_net(np.load(path, allow_pickle=True)[()])
