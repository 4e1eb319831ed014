"""A minimal stand-in for `h5py` (absent from this image), test infrastructure only.

It implements the part of the h5py surface the reference's data/vr_bag.py and the
product's vr_bag.py touch: `File(path, mode)` as a context manager, groups indexed by
name with `in`, datasets with `.shape`, `.attrs`, NumPy-style `__getitem__` (always a
copy, like h5py), `__setitem__` (slices, `[:]`, and the `[0, i, 'field']` spelling),
and `[()]`.  The container is an .npz written through a file object (so the `.bag`
suffix is kept and `shutil.copy` works on it); `r+` files are written back on close.

    import tests.fake_h5py as fake; fake.install()      # registers sys.modules['h5py']
"""
import io
import os
import sys
import types

import numpy as np


class Dataset:
    def __init__(self, owner, arr, attrs):
        self._owner, self._arr, self.attrs = owner, arr, attrs

    shape = property(lambda self: self._arr.shape)
    dtype = property(lambda self: self._arr.dtype)

    @staticmethod
    def _split(key):
        if isinstance(key, tuple) and key and isinstance(key[-1], str):
            return key[:-1], key[-1]
        return key, None

    def __getitem__(self, key):
        key, field = self._split(key)
        out = self._arr[key]
        if field is not None:
            out = out[field]
        return np.array(out, copy=True) if isinstance(out, np.ndarray) else out

    def __setitem__(self, key, value):
        if self._owner.mode == "r":
            raise OSError("file is open read-only")
        key, field = self._split(key)
        if field is not None:
            self._arr[field][key] = value
        else:
            self._arr[key] = value
        self._owner._dirty = True


class Group:
    def __init__(self, owner, prefix):
        self._owner, self._prefix = owner, prefix

    def _children(self):
        p = self._prefix
        return {k[len(p):].split("/")[0] for k in self._owner._data if k.startswith(p)}

    def __contains__(self, name):
        return name in self._children()

    def keys(self):
        return sorted(self._children())

    def __getitem__(self, name):
        full = self._prefix + name
        if full in self._owner._data:
            return Dataset(self._owner, self._owner._data[full], self._owner._attrs.setdefault(full, {}))
        if name in self._children():
            return Group(self._owner, full + "/")
        raise KeyError(name)


class File(Group):
    def __init__(self, path, mode="r"):
        self.path, self.mode, self._dirty = str(path), mode, False
        self._data, self._attrs = {}, {}
        if mode in ("r", "r+", "a") and os.path.exists(self.path):
            with open(self.path, "rb") as f:
                z = np.load(io.BytesIO(f.read()), allow_pickle=True)
            for k in z.files:
                if k.startswith("__attrs__/"):
                    self._attrs[k[len("__attrs__/"):]] = z[k].item()
                else:
                    self._data[k] = z[k]
        elif mode in ("r", "r+"):
            raise FileNotFoundError(self.path)
        super().__init__(self, "")

    def create_dataset(self, name, data, attrs=None):
        self._data[name] = np.array(data)
        if attrs:
            self._attrs[name] = dict(attrs)
        self._dirty = True

    def close(self):
        if self._dirty and self.mode != "r":
            blob = dict(self._data)
            for k, v in self._attrs.items():
                if v:
                    blob["__attrs__/" + k] = np.array(v, dtype=object)
            with open(self.path, "wb") as f:
                np.savez(f, **blob)
            self._dirty = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def install():
    mod = types.ModuleType("h5py")
    mod.File, mod.Dataset, mod.Group = File, Dataset, Group
    mod.__version__ = "0.0-fake"
    sys.modules["h5py"] = mod
    return mod


# ---------------------------------------------------------------- synthetic containers
VR_META_DTYPE = np.dtype([("index", np.uint32), ("dimensions_x", np.uint32), ("dimensions_y", np.uint32),
                          ("resolution_x", np.float32), ("resolution_y", np.float32),
                          ("sw_corner_x", np.float32), ("sw_corner_y", np.float32)])
VR_REF_DTYPE = np.dtype([("depth", np.float32), ("depth_uncrt", np.float32)])


def write_vr_bag(path, base_shape=(5, 7), seed=3, refined_frac=0.7, lo=2, hi=14, shuffle_storage=True, nodata_frac=0.12):
    """A synthetic VR BAG: `base_shape` base cells, a random subset refined with grids of lo..hi cells a side and
    0.5 / 1 / 2 / 4 m resolution; the record blocks are stored in a SHUFFLED order when `shuffle_storage` (the
    `index` field is what locates a grid, not its position in the base grid).  One refined cell is fully nodata."""
    rng = np.random.default_rng(seed)
    R, C = base_shape
    meta = np.zeros((R, C), VR_META_DTYPE)
    meta["index"] = 4294967295
    cells = [(r, c) for r in range(R) for c in range(C) if rng.random() < refined_frac]
    order = list(range(len(cells)))
    if shuffle_storage:
        rng.shuffle(order)
    dims = {i: (int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))) for i in range(len(cells))}
    recs, start, pos = [None] * len(cells), {}, 0
    for i in order:
        ny, nx = dims[i]
        start[i] = pos
        pos += ny * nx
    ref = np.zeros((1, pos), VR_REF_DTYPE)
    for i, (r, c) in enumerate(cells):
        ny, nx = dims[i]
        res = float(rng.choice([0.5, 1.0, 2.0, 4.0]))
        d = (25.0 + 0.04 * np.arange(nx)[None, :] + 0.02 * np.arange(ny)[:, None] + rng.normal(0, 0.2, (ny, nx))).astype(np.float32)
        d = (d + (rng.random((ny, nx)) < 0.03) * rng.uniform(1, 4, (ny, nx))).astype(np.float32)
        d[rng.random((ny, nx)) < nodata_frac] = 1.0e6
        if i == 2:
            d[:] = 1.0e6
        if i == 4:
            d[0, 0] = np.nan
        u = (0.1 + 0.05 * rng.random((ny, nx))).astype(np.float32)
        a = start[i]
        ref["depth"][0, a:a + ny * nx] = d.reshape(-1)
        ref["depth_uncrt"][0, a:a + ny * nx] = u.reshape(-1)
        meta[r, c] = (a, nx, ny, res, res, float(rng.integers(0, 3)) * res, float(rng.integers(0, 3)) * res)
    f = File(path, "w")
    f.create_dataset("BAG_root/elevation", rng.normal(30, 2, (R, C)).astype(np.float32),
                     attrs={"Minimum Elevation Value": 20.0, "Maximum Elevation Value": 40.0})
    f.create_dataset("BAG_root/uncertainty", np.full((R, C), 0.5, np.float32))
    f.create_dataset("BAG_root/varres_metadata", meta)
    f.create_dataset("BAG_root/varres_refinements", ref)
    f.create_dataset("BAG_root/metadata", np.frombuffer(b"<xml/>", dtype="S1"))
    f.close()
    return meta, ref


def write_sr_bag(path, shape=(23, 31), seed=4, resolution=2.0, with_uncertainty=True):
    rng = np.random.default_rng(seed)
    d = (30.0 + rng.normal(0, 0.3, shape)).astype(np.float32)
    d[rng.random(shape) < 0.1] = 1.0e6
    xml = (f"<gmd:MD_Metadata><gmd:resolution><gco:Measure uom=\"m\">{resolution}</gco:Measure>"
           "</gmd:resolution></gmd:MD_Metadata>").encode()
    f = File(path, "w")
    f.create_dataset("BAG_root/elevation", d)
    if with_uncertainty:
        f.create_dataset("BAG_root/uncertainty", (0.2 + 0.1 * rng.random(shape)).astype(np.float32))
    f.create_dataset("BAG_root/metadata", np.frombuffer(xml, dtype="S1"))
    f.close()
    return d
