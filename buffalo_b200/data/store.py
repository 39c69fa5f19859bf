"""Single-file array store with the slice of the h5py API the data layer touches:
File(path, "r"|"w"), attrs, create_group, create_dataset, `name in f`, keys(), dataset slicing / assignment,
.shape, .chunks, close().  h5py is not installed in this image; the on-disk format is one uncompressed .npz
written to exactly `path` (caching is keyed on os.path.isfile(opt.data.path), buffalo/data/mm.py:241-245)."""
import json
import os

import numpy as np


class Dataset(object):
    def __init__(self, arr):
        self._a = arr

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)
    chunks = property(lambda self: (max(1, min(len(self._a), 1 << 24)),))

    def __getitem__(self, idx):
        return self._a[idx]

    def __setitem__(self, idx, value):
        if self._a.dtype.kind == "S":
            value = np.asarray([v.encode("utf-8") if isinstance(v, str) else v for v in np.atleast_1d(value)])
        self._a[idx] = value

    def __len__(self):
        return len(self._a)

    def __iter__(self):
        return iter(self._a)

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)


class Group(object):
    def __init__(self):
        self._items = {}
        self.attrs = {}

    def create_group(self, name):
        g = Group()
        self._items[name] = g
        return g

    def create_dataset(self, name, shape=None, dtype="float32", data=None, **unused):
        arr = np.asarray(data) if data is not None else np.zeros(shape, dtype=dtype)
        ds = Dataset(arr)
        self._items[name] = ds
        return ds

    def __getitem__(self, name):
        return self._items[name]

    def __contains__(self, name):
        return name in self._items

    def keys(self):
        return self._items.keys()


class File(Group):
    def __init__(self, path, mode="r"):
        super().__init__()
        self.path, self.mode = path, mode
        if mode == "r":
            self._load()

    def _walk(self, group, prefix, arrays, attrs):
        attrs[prefix] = {k: (v.item() if isinstance(v, np.generic) else v) for k, v in group.attrs.items()}
        for name, item in group._items.items():
            key = prefix + "/" + name if prefix else name
            if isinstance(item, Group):
                self._walk(item, key, arrays, attrs)
            else:
                arrays[key] = item._a

    def close(self):
        if self.mode == "w":
            arrays, attrs = {}, {}
            self._walk(self, "", arrays, attrs)
            arrays["__meta__"] = np.frombuffer(json.dumps(attrs).encode("utf-8"), dtype=np.uint8)
            tmp = self.path + ".tmp"
            with open(tmp, "wb") as fout:      # straight to the file: no in-memory copy of a billion-entry database
                np.savez(fout, **arrays)
            os.replace(tmp, self.path)
            self.mode = "closed"

    def _load(self):
        with np.load(self.path, allow_pickle=False) as z:
            attrs = json.loads(bytes(z["__meta__"]).decode("utf-8"))
            for prefix, a in attrs.items():
                node = self
                for part in [p for p in prefix.split("/") if p]:
                    node = node._items[part] if part in node._items else node.create_group(part)
                node.attrs.update(a)
            for key in z.files:
                if key == "__meta__":
                    continue
                parts = key.split("/")
                node = self
                for part in parts[:-1]:
                    node = node._items[part] if part in node._items else node.create_group(part)
                node._items[parts[-1]] = Dataset(z[key])
