"""TEST INFRASTRUCTURE -- a minimal stand-in for h5py (which is not in the image), enough of its
API for the HDF5 readers / writers of this package AND of the reference (results.py:987-1158,
response.py:578-635, 782-802): File / Group with .attrs, create_group, create_dataset, item
access, membership, iteration; datasets indexed with [...] and [()].  A "file" is a pickle of
the tree, so what one side writes the other side reads through the same calls a real HDF5 file
would see.  install() registers it as `h5py` (also onto an already-registered stub module)."""
import pickle
import sys
import types

import numpy as np


class Dataset(object):
    def __init__(self, data):
        self._a = np.array(data)

    @property
    def shape(self):
        return self._a.shape

    def __getitem__(self, key):
        if key is Ellipsis:
            return self._a.copy()
        if key == ():
            return self._a[()]
        return self._a[key]


class Group(object):
    def __init__(self):
        self.attrs = {}
        self._items = {}

    def create_group(self, name):
        g = self._items[name] = Group()
        return g

    def create_dataset(self, name, data=None):
        d = self._items[name] = Dataset(data)
        return d

    def __getitem__(self, name):
        return self._items[name]

    def __contains__(self, name):
        return name in self._items

    def __iter__(self):
        return iter(self._items)

    def keys(self):
        return self._items.keys()


def _dump(g):
    return {"attrs": {k: np.array(v) if isinstance(v, (list, tuple)) else v for k, v in g.attrs.items()},
            "items": {k: (_dump(v) if isinstance(v, Group) else v._a) for k, v in g._items.items()}}


def _restore(g, d):
    g.attrs = dict(d["attrs"])
    for k, v in d["items"].items():
        if isinstance(v, dict):
            _restore(g.create_group(k), v)
        else:
            g.create_dataset(k, data=v)


class File(Group):
    def __init__(self, filename, mode="r"):
        Group.__init__(self)
        self._filename, self._mode = filename, mode
        if mode == "r":
            with open(filename, "rb") as fh:
                _restore(self, pickle.load(fh))

    def close(self):
        if self._mode != "r":
            with open(self._filename, "wb") as fh:
                pickle.dump(_dump(self), fh)


def install():
    mod = sys.modules.get("h5py")
    if mod is None or not isinstance(mod, types.ModuleType):
        mod = sys.modules["h5py"] = types.ModuleType("h5py")
    mod.File, mod.Group, mod.Dataset = File, Group, Dataset
    mod._hl = types.SimpleNamespace(group=types.SimpleNamespace(Group=Group))
    mod.__fake__ = True
    return mod


def uninstall():
    mod = sys.modules.get("h5py")
    if mod is not None and getattr(mod, "__fake__", False):
        del sys.modules["h5py"]
