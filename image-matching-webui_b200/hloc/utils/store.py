"""Feature / match container with the layout of the reference's HDF5 files (imcui/hloc/extract_features.py:223-241,
match_features.py:73-83): one GROUP per image name (or per `names_to_pair` key) holding named DATASETS with the reference's
dtypes (keypoints / descriptors / scores fp16 when as_half, image_size int, matches0 int16, matching_scores0 fp16) and
dataset ATTRIBUTES (keypoints.uncertainty).

Two back ends behind one interface:
  * `.h5` paths use h5py (the reference's on-disk format, readable by hloc's SfM pipelines) -- when h5py is installed;
  * any other suffix uses the built-in append-only record file below (this image has no h5py): records
    `[u32 header bytes][JSON {group, name, dtype, shape, attrs}][raw array bytes]`, later records of the same
    (group, name) supersede earlier ones, the index is rebuilt by one sequential scan on open.
"""
import json
import struct
import threading
from pathlib import Path

import numpy as np

try:  # pragma: no cover - not installed in the build image
    import h5py
except Exception:  # noqa: BLE001
    h5py = None


class RecordFile:
    """Append-only group/dataset store (see module docstring).  Thread-safe for one writer per process."""
    MAGIC = b"IMWSTORE1\n"

    def __init__(self, path, mode="a"):
        self.path, self.mode = Path(path), mode
        self.index = {}          # (group, name) -> (offset, dtype, shape, attrs)
        self.lock = threading.Lock()
        exists = self.path.exists() and self.path.stat().st_size > 0
        if mode == "r" and not exists:
            raise FileNotFoundError(str(self.path))
        if mode == "w" or not exists:
            self.path.parent.mkdir(parents=True, exist_ok=True)
            with open(self.path, "wb") as f:
                f.write(self.MAGIC)
        self._scan()
        self.f = open(self.path, "rb" if mode == "r" else "r+b")

    def _scan(self):
        with open(self.path, "rb") as f:
            if f.read(len(self.MAGIC)) != self.MAGIC:
                raise ValueError(f"{self.path}: not an imw record file")
            while True:
                raw = f.read(4)
                if len(raw) < 4:
                    break
                (n,) = struct.unpack("<I", raw)
                head = json.loads(f.read(n))
                nbytes = int(np.dtype(head["dtype"]).itemsize * int(np.prod(head["shape"], dtype=np.int64)))
                off = f.tell()
                if head.get("deleted"):
                    for k in [k for k in self.index if k[0] == head["group"]]:
                        del self.index[k]
                else:
                    self.index[(head["group"], head["name"])] = (off, head["dtype"], tuple(head["shape"]), head.get("attrs", {}))
                f.seek(off + nbytes)

    # ---- write ----
    def write_group(self, group, arrays, attrs=None, replace=True):
        """arrays: {dataset name: ndarray}; attrs: {dataset name: {attr: value}}.  `replace` drops the group's older datasets."""
        assert self.mode != "r"
        with self.lock:
            self.f.seek(0, 2)
            if replace and any(k[0] == group for k in self.index):
                head = json.dumps({"group": group, "name": "", "dtype": "u1", "shape": [0], "deleted": True}).encode()
                self.f.write(struct.pack("<I", len(head)) + head)
                for k in [k for k in self.index if k[0] == group]:
                    del self.index[k]
            for name, a in arrays.items():
                a = np.ascontiguousarray(a)
                at = {k: (float(v) if np.isscalar(v) or getattr(v, "ndim", 1) == 0 else np.asarray(v).tolist())
                      for k, v in (attrs or {}).get(name, {}).items()}
                head = json.dumps({"group": group, "name": name, "dtype": a.dtype.str, "shape": list(a.shape), "attrs": at}).encode()
                self.f.write(struct.pack("<I", len(head)) + head)
                off = self.f.tell()
                self.f.write(a.tobytes())
                self.index[(group, name)] = (off, a.dtype.str, tuple(a.shape), at)
            self.f.flush()

    # ---- read ----
    def groups(self):
        return sorted({g for g, _ in self.index})

    def __contains__(self, group):
        return any(g == group for g, _ in self.index)

    def datasets(self, group):
        return [n for g, n in self.index if g == group]

    def read(self, group, name):
        off, dt, shape, _ = self.index[(group, name)]
        with self.lock:
            self.f.seek(off)
            raw = self.f.read(int(np.dtype(dt).itemsize * int(np.prod(shape, dtype=np.int64))))
        return np.frombuffer(raw, dtype=np.dtype(dt)).reshape(shape).copy()

    def attrs(self, group, name):
        return dict(self.index[(group, name)][3])

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class H5File:  # pragma: no cover - exercised only where h5py exists
    """The same interface over h5py (`libver="latest"`, as the reference opens its files)."""

    def __init__(self, path, mode="a"):
        if h5py is None:
            raise ImportError("h5py is not installed: use a non-.h5 suffix for the built-in record file")
        self.f = h5py.File(str(path), mode, libver="latest")

    def write_group(self, group, arrays, attrs=None, replace=True):
        if replace and group in self.f:
            del self.f[group]
        grp = self.f.require_group(group)
        for name, a in arrays.items():
            if name in grp:
                del grp[name]
            grp.create_dataset(name, data=a)
            for k, v in (attrs or {}).get(name, {}).items():
                grp[name].attrs[k] = v

    def groups(self):
        names = []
        self.f.visititems(lambda _, obj: names.append(obj.parent.name.strip("/")) if isinstance(obj, h5py.Dataset) else None)
        return sorted(set(names))

    def __contains__(self, group):
        return group in self.f

    def datasets(self, group):
        return list(self.f[group].keys())

    def read(self, group, name):
        return self.f[group][name].__array__()

    def attrs(self, group, name):
        return dict(self.f[group][name].attrs)

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open_store(path, mode="a"):
    return H5File(path, mode) if str(path).endswith((".h5", ".hdf5")) else RecordFile(path, mode)
