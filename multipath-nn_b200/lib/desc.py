"""`net_desc` statistics and their text rendering -- same nested-dict schema
and log format as /root/reference/scripts/lib/desc.py:10-79.
The per-batch tensors come from `net.eval_stats` (GPU) instead of a TF session;
the dataset reduction (sum over examples / count, in float64 on the host) is
the reference's.
"""
import numpy as np

__all__ = ['state_tensors', 'mean_net_state', 'net_desc', 'render_net_desc']


def state_tensors(net):
    """Keys of the statistics `eval_stats` produces (train-nets:111-130)."""
    keys = [(net, 'acc'), (net, 'moc')]
    for l in net.leaves:
        keys += [(l, 'p_cor'), (l, 'p_inc'), (l, 'p_cor_by_cls'), (l, 'p_inc_by_cls')]
        if net.dynamic:
            keys.append((l, 'p_tr'))
        keys.append((l, 'c_err'))
    for l in net.layers:
        if l.router is not None:
            keys.append((l, 'x_rte'))
    return {k: k for k in keys}


def mean_net_state(net, tensors, data, hypers):
    if len(tensors) == 0:
        return {}
    sums = {k: 0 for k in tensors.keys()}
    count = 0
    for x0, y in data:
        samples = net.eval_stats({net.x0: x0, net.y: y, **hypers})
        for k in tensors.keys():
            sums[k] = sums[k] + np.sum(np.asarray(samples[k], dtype=np.float64), 0)
        count += len(x0)
    return {k: (sums[k] / count).tolist() for k in tensors.keys()}


def layer_desc(layer, stats_tr, stats_ts):
    return {'name': layer.name,
            'stats_tr': {k: v for (t, k), v in stats_tr.items() if t is layer},
            'stats_ts': {k: v for (t, k), v in stats_ts.items() if t is layer},
            'sinks': [layer_desc(s, stats_tr, stats_ts) for s in layer.sinks]}


def net_desc(net, dataset, hypers={}, state={}):
    stats_tr = mean_net_state(net, state, dataset.training_set(), hypers)
    stats_ts = mean_net_state(net, state, dataset.test_set(), hypers)
    return {'type': type(net).__name__,
            'stats_tr': {k: v for (t, k), v in stats_tr.items() if t is net},
            'stats_ts': {k: v for (t, k), v in stats_ts.items() if t is net},
            'root': layer_desc(net.root, stats_tr, stats_ts)}


def render_stats(stats):
    if len(stats) == 0:
        return ''
    scalars = [kv for kv in sorted(stats.items()) if np.ndim(kv[1]) == 0]
    return '(%s)' % '; '.join('%s=%.3g' % kv for kv in scalars)


def render_layer_desc(desc, stats_key):
    lines = '%s %s' % (desc['name'], render_stats(desc[stats_key]))
    n = len(desc['sinks'])
    for i, s in enumerate(desc['sinks']):
        cont = '\n| ' if i < n - 1 else '\n  '
        lines += '\n↳ ' + render_layer_desc(s, stats_key).replace('\n', cont)
    return lines


def render_net_desc(desc, name='Network'):
    bar = '─' * 59
    ind = '\n│     '
    body = [
        '┌' + bar, '│ ' + name, '├' + bar,
        '│ Training Set:', '│',
        '│   [%s] %s' % (desc['type'], render_stats(desc['stats_tr'])),
        '│     ' + render_layer_desc(desc['root'], 'stats_tr').replace('\n', ind),
        '│', '│ Test Set:', '│',
        '│   [%s] %s' % (desc['type'], render_stats(desc['stats_ts'])),
        '│     ' + render_layer_desc(desc['root'], 'stats_ts').replace('\n', ind),
        '│']
    return '\n'.join(body)
