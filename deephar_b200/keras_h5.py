"""Keras 2.1.4 HDF5 weight files <-> deephar_b200 weight tables.

File structure (keras/engine/topology.py::save_weights_to_hdf5_group, restated):
  /                         attrs: layer_names = [b'<layer>', ...], backend, keras_version
  /<layer>                  attrs: weight_names = [b'<tf variable name>:0', ...]
  /<layer>/<weight name>    dataset (the weight name contains '/', so h5py nests groups)
A file written by `model.save()` holds the same tree under /model_weights.  `<layer>` is a
top-level layer of the saved model: a plain layer (`conv1`, `batch_normalization_7`), a
TimeDistributed wrapper, or a nested sub-model (`Stem`, `rBlock3`, `SepConv3`, `RegMap3`,
`fReMap3`, ... -- reception.py:92,129,140,151,162) whose group then lists the weights of all
its inner layers: trainable ones first, then the BatchNormalization moving statistics.

Names on this side are "<sub-model>/<layer>/<leaf>" or "<layer>/<leaf>" (Model.weight_specs).
A file entry (group G, weight name W) is mapped by trying, in order,
  G/W            nested sub-model, variables named '<inner layer>/<leaf>:0'
  W              variables already carrying the full scope
  G/<leaf of W>  wrapper layers whose variable scope differs from the layer name
with ':<n>' suffixes dropped and the TimeDistributed prefix 'td_' that the CVPR'18 merge model
puts around the pose network's sub-models (action.py:117-153) stripped.

Entries of the file that match nothing are the constants the reference assigns itself
(soft-argmax grids, aggregation matrix: layers.py:160-200, blocks.py:229-233) or layers the
target model does not have; they are returned as `unused`, never an error.
"""
import numpy as np

from . import hdf5

STRIP_PREFIXES = ('td_',)


def _as_str(b):
    return b.decode('utf-8') if isinstance(b, (bytes, np.bytes_)) else str(b)


def _weights_root(f):
    if 'layer_names' not in f.attrs and 'model_weights' in f.keys():
        return f['model_weights']
    return f


def read_entries(path):
    """-> ([(group, weight_name, ndarray)], file attrs) in the file's layer / weight order.  A file that is not a
    readable Keras weight file (truncated download, damaged bytes, a layer group that misses a listed weight) raises
    hdf5.Hdf5Error, never a parser-internal exception."""
    try:
        return _read_entries(path)
    except hdf5.Hdf5Error:
        raise
    except hdf5.CORRUPT as e:
        raise hdf5.Hdf5Error('%s: corrupt or truncated Keras HDF5 weight file (%s: %s)' % (path, type(e).__name__, e))


def _read_entries(path):
    with hdf5.File(path) as f:
        root = _weights_root(f)
        if 'layer_names' in root.attrs:
            layers = [_as_str(n) for n in np.atleast_1d(root.attrs['layer_names'])]
        else:                       # not written by Keras: take the groups as they come
            layers = root.keys()
        entries = []
        for lname in layers:
            g = root[lname]
            if 'weight_names' in g.attrs:
                wnames = [_as_str(n) for n in np.atleast_1d(g.attrs['weight_names'])]
            else:
                wnames = [p for p, _ in g.visit_datasets()]
            for wn in wnames:
                entries.append((lname, wn, np.asarray(g[wn].read())))
        attrs = {k: v for k, v in root.attrs.items() if k not in ('layer_names',)}
    return entries, attrs


def _strip(name):
    for pre in STRIP_PREFIXES:
        if name.startswith(pre):
            return name[len(pre):]
    return name


def candidates(group, wname):
    parts = wname.split('/')
    parts[-1] = parts[-1].split(':')[0]
    parts = [_strip(p) for p in parts]
    g = _strip(group)
    w = '/'.join(parts)
    out = [g + '/' + w, w, g + '/' + parts[-1]]
    if len(parts) >= 2:
        out.append('/'.join(parts[-2:]))
    seen, uniq = set(), []
    for c in out:
        if c not in seen:
            seen.add(c)
            uniq.append(c)
    return uniq


def match_entries(entries, weight_specs):
    """-> ({model weight name: array}, [unused (group, weight_name)]).  Shape mismatches raise."""
    specs = dict(weight_specs)
    table, unused = {}, []
    for group, wname, arr in entries:
        hit = None
        for c in candidates(group, wname):
            if c in specs and c not in table:
                hit = c
                break
        if hit is None:
            unused.append((group, wname))
            continue
        if tuple(arr.shape) != tuple(specs[hit]):
            raise ValueError('%s/%s has shape %s in the file, the model expects %s for %s'
                             % (group, wname, tuple(arr.shape), tuple(specs[hit]), hit))
        table[hit] = arr.astype(np.float32, copy=False)
    return table, unused


def load(path, weight_specs, optional=(), by_name=False):
    """Weights of a Keras .h5 for a model with `weight_specs`.
    by_name=False (keras topological loading: the file was saved from the same architecture):
    every non-optional model weight must be present.  by_name=True: the intersection is
    returned and the caller keeps its current values for the rest (keras semantics)."""
    entries, _ = read_entries(path)
    table, unused = match_entries(entries, weight_specs)
    if not by_name:
        missing = [n for n, _ in weight_specs if n not in table and n not in set(optional)]
        if missing:
            raise KeyError('%s: %d model weights are not in the file (first: %s); use by_name=True to load '
                           'a partial checkpoint' % (path, len(missing), ', '.join(missing[:4])))
    return table, unused


def save(path, weight_specs, table, backend='tensorflow', keras_version='2.1.4'):
    """Write `table` the way keras.Model.save_weights would for the same architecture: one group per
    top-level layer (sub-model scopes become nested-model groups), trainable weights before the
    BatchNormalization moving statistics inside a group."""
    groups, order = {}, []
    for name, _ in weight_specs:
        if name not in table:
            continue
        parts = name.split('/')
        g = parts[0]
        inner = '/'.join(parts[1:]) if len(parts) > 2 else name
        if g not in groups:
            groups[g] = []
            order.append(g)
        groups[g].append((inner + ':0', np.asarray(table[name], dtype=np.float32)))
    with hdf5.Writer(path) as w:
        w.set_attr('/', 'layer_names', np.array([g.encode('utf-8') for g in order]))
        w.set_attr('/', 'backend', backend)
        w.set_attr('/', 'keras_version', keras_version)
        for g in order:
            ws = groups[g]
            moving = ('moving_mean:0', 'moving_variance:0')
            ws = [e for e in ws if not e[0].endswith(moving)] + [e for e in ws if e[0].endswith(moving)]
            w.create_group(g)
            w.set_attr(g, 'weight_names', np.array([n.encode('utf-8') for n, _ in ws]))
            for n, a in ws:
                w.create_dataset(g + '/' + n, a)
