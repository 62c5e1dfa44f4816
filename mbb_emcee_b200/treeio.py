"""Results / response files in the reference's HDF5 layout, on two carriers.

The reference serialises fits and passband responses with h5py
(mbb_emcee/results.py:987-1158, response.py:578-635, 782-802): a tree of groups,
each with attributes and datasets.  Here the objects produce / consume that tree
as plain nested dicts

    {"attrs": {name: value}, "data": {name: ndarray}, "groups": {name: tree}}

with the reference's group, attribute and dataset names, and this module moves a
tree to and from a file:

  * ``.h5`` / ``.hdf5`` -- a real HDF5 file through h5py, readable by the
    reference (and the reference's files readable here).  h5py is not a
    dependency of this package; ImportError when it is missing.
  * anything else -- a ``.npz`` archive: dataset ``D`` of group ``A/B`` is the
    entry ``A/B/D``, attribute ``X`` of that group the entry ``A/B/@X`` (root
    attributes ``@X``).  No extra dependency.
"""
import numpy as np

__all__ = ["new_tree", "write_tree", "read_tree", "is_hdf5_name"]


def new_tree():
    return {"attrs": {}, "data": {}, "groups": {}}


def is_hdf5_name(filename):
    return str(filename).lower().endswith((".h5", ".hdf5", ".hdf"))


def npz_name(filename):
    """np.savez appends '.npz' to names without it: say so up front instead."""
    filename = str(filename)
    return filename if filename.lower().endswith(".npz") else filename + ".npz"


# ------------------------------------------------------------------------- h5py
def _h5_write(handle, tree):
    for k, v in tree["attrs"].items():
        handle.attrs[k] = v
    for k, v in tree["data"].items():
        handle.create_dataset(k, data=v)
    for k, sub in tree["groups"].items():
        _h5_write(handle.create_group(k), sub)


def _h5_read(handle):
    tree = new_tree()
    for k in handle.attrs:
        tree["attrs"][k] = handle.attrs[k]
    for k in handle:
        item = handle[k]
        if hasattr(item, "create_group") or (hasattr(item, "keys") and not hasattr(item, "shape")):
            tree["groups"][k] = _h5_read(item)
        else:
            tree["data"][k] = item[...] if getattr(item, "shape", ()) != () else item[()]
    return tree


def _import_h5py():
    try:
        import h5py
    except ImportError:
        raise ImportError("HDF5 files need h5py, which is not installed; use a '.npz' file name "
                          "(same groups, attributes and datasets) instead")
    return h5py


# -------------------------------------------------------------------------- npz
def _npz_flatten(tree, prefix, out):
    for k, v in tree["attrs"].items():
        out[prefix + "@" + k] = np.asarray(v)
    for k, v in tree["data"].items():
        out[prefix + k] = np.asarray(v)
    for k, sub in tree["groups"].items():
        if "/" in k or k.startswith("@"):
            raise ValueError("group name %r cannot be stored" % (k,))
        _npz_flatten(sub, prefix + k + "/", out)


def _npz_unflatten(f):
    tree = new_tree()
    for key in f.files:
        parts = key.split("/")
        node = tree
        for g in parts[:-1]:
            node = node["groups"].setdefault(g, new_tree())
        leaf, val = parts[-1], f[key]
        if val.dtype.kind in "US" and val.shape == ():
            val = str(val[()])
        elif val.shape == ():
            val = val[()]
        if leaf.startswith("@"):
            node["attrs"][leaf[1:]] = val
        else:
            node["data"][leaf] = val
    return tree


# ---------------------------------------------------------------------- public
def write_tree(filename, tree):
    """Returns the name of the file written."""
    if is_hdf5_name(filename):
        h5py = _import_h5py()
        f = h5py.File(filename, "w")
        try:
            _h5_write(f, tree)
        finally:
            f.close()
        return str(filename)
    out = {}
    _npz_flatten(tree, "", out)
    name = npz_name(filename)
    with open(name, "wb") as fh:           # a file object: numpy then leaves the name alone
        np.savez_compressed(fh, **out)
    return name


def read_tree(filename):
    if is_hdf5_name(filename):
        h5py = _import_h5py()
        f = h5py.File(filename, "r")
        try:
            return _h5_read(f)
        finally:
            f.close()
    with np.load(filename, allow_pickle=False) as f:
        return _npz_unflatten(f)
