"""Parameter files in the reference's on-disk format (the format
is defined by /root/reference/scripts/lib/serdes.py:13-60).

A file is `np.save` of one pickled, nested record:

    net record   = {'type', 'root': layer record, 'hypers': {...}, 'params': {name: ndarray}}
    layer record = {'type', 'name', 'hypers': {...}, 'params': {name: float32 ndarray},
                    'sinks': [layer record], 'comps': [layer record], 'router': layer record | None}

Class names, hyper-parameter keys (including the unicode ones) and parameter
keys are the reference's, so files written by either implementation load in the
other.  NumPy >= 1.16.3 needs `allow_pickle=True` on load, which the reference's
`np.load(path)[()]` predates.

The codec below is table-driven: `_CHILDREN` names the three places a layer
record can hold other layer records, and one walker pairs a linked layer tree
with its record for the parameter transfer.
"""
import numpy as np

import lib.layer_types
import lib.net_types

__all__ = ['encode_layer', 'decode_layer', 'load_params', 'encode_net', 'decode_net',
           'write_net', 'read_net']

# (record key, holds a list?) for every slot of a layer that contains layers
_CHILDREN = (('sinks', True), ('comps', True), ('router', False))


def _arrays(namespace):
    """current value of every parameter of a layer / net, by name"""
    return {key: param.eval() for key, param in vars(namespace).items()}


def _assign(namespace, arrays):
    for key, value in arrays.items():
        getattr(namespace, key).assign(value)


def encode_layer(layer):
    """layer tree -> record (None stays None: 'no router')"""
    if layer is None:
        return None
    record = {'type': type(layer).__name__, 'name': layer.name,
              'hypers': dict(vars(layer.hypers)), 'params': _arrays(layer.params)}
    for slot, many in _CHILDREN:
        held = getattr(layer, slot)
        record[slot] = [encode_layer(child) for child in held] if many else encode_layer(held)
    return record


def decode_layer(record):
    """record -> unlinked layer tree of the same classes and hypers (parameters come later:
    they only exist once the net has been linked, see `load_params`)"""
    if record is None:
        return None
    kwargs = dict(record['hypers'])
    for slot, many in _CHILDREN:
        held = record[slot]
        kwargs[slot] = [decode_layer(r) for r in held] if many else decode_layer(held)
    return getattr(lib.layer_types, record['type'])(name=record['name'], **kwargs)


def _paired(layer, record):
    """(layer, record) for a linked tree and its record, children matched by position"""
    todo = [(layer, record)]
    while todo:
        node, rec = todo.pop()
        if node is None:
            continue
        yield node, rec
        for slot, many in _CHILDREN:
            if many:
                todo.extend(zip(getattr(node, slot), rec[slot]))
            else:
                todo.append((getattr(node, slot), rec[slot]))


def load_params(layer, record):
    """copy every stored array into the linked layer tree"""
    for node, rec in _paired(layer, record):
        _assign(node.params, rec['params'])


def encode_net(net):
    eng = getattr(net, '_engine', None)
    if eng is not None and not getattr(eng, 'dry', False) and not eng._snapshot:
        with eng.host_snapshot():                  # one device-to-host copy instead of one per tensor
            return encode_net(net)
    return {'type': type(net).__name__, 'root': encode_layer(net.root),
            'hypers': dict(vars(net.hypers)), 'params': _arrays(net.params)}


def decode_net(record):
    make = getattr(lib.net_types, record['type'])
    net = make(root=decode_layer(record['root']), **dict(record['hypers']))     # links the tree
    load_params(net.root, record['root'])
    _assign(net.params, record['params'])
    return net


def write_net(path, net):
    np.save(path, encode_net(net))


def read_net(path):
    return decode_net(np.load(path, allow_pickle=True)[()])
