"""PCD files (the format of the reference's real-camera captures, test.py:102): fields x y z and optionally a packed
rgb / rgba column (float32 or uint32 holding 0x00RRGGBB), DATA ascii or binary."""
import numpy as np

from .geometry import PointCloud


def _header(f):
    h = {}
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PCD: no DATA line")
        parts = line.decode("ascii", "replace").strip().split()
        if not parts or parts[0].startswith("#"):
            continue
        h[parts[0].upper()] = parts[1:]
        if parts[0].upper() == "DATA":
            return h


def read_point_cloud(filename, *args, **kwargs):
    with open(filename, "rb") as f:
        h = _header(f)
        fields, sizes, types = h["FIELDS"], [int(s) for s in h["SIZE"]], h["TYPE"]
        counts = [int(c) for c in h.get("COUNT", ["1"] * len(fields))]
        n = int(h["POINTS"][0]) if "POINTS" in h else int(h["WIDTH"][0]) * int(h["HEIGHT"][0])
        kind = h["DATA"][0].lower()
        np_type = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4", ("I", 1): "i1",
                   ("I", 2): "<i2", ("I", 4): "<i4"}
        dtype = np.dtype([(name, np_type[(t, s)], (c,)) if c > 1 else (name, np_type[(t, s)])
                          for name, s, t, c in zip(fields, sizes, types, counts)])
        if kind == "binary":
            data = np.frombuffer(f.read(n * dtype.itemsize), dtype=dtype, count=n)
        elif kind == "ascii":
            rows = np.loadtxt(f, dtype=np.float64, ndmin=2)[:n]
            data = np.zeros(len(rows), dtype=dtype)
            for i, name in enumerate(fields):
                data[name] = rows[:, i]          # cast to the declared type (a float-typed rgb is re-read bitwise below)
        else:
            raise ValueError(f"PCD: DATA {kind} is not supported by this stand-in")
    cloud = PointCloud()
    xyz = np.stack([data["x"], data["y"], data["z"]], axis=1).astype(np.float64)
    ok = np.isfinite(xyz).all(1)
    cloud.points = xyz[ok]
    for name in ("rgb", "rgba"):
        if name in fields:
            packed = np.ascontiguousarray(data[name][ok])
            packed = packed.view(np.uint32) if packed.dtype.kind == "f" else packed.astype(np.uint32)
            cloud.colors = np.stack([(packed >> 16) & 255, (packed >> 8) & 255, packed & 255], axis=1) / 255.0
    return cloud


def write_point_cloud(filename, pointcloud, write_ascii=False, *args, **kwargs):
    pts = np.asarray(pointcloud.points, dtype=np.float32)
    has_rgb = pointcloud.has_colors()
    fields = "x y z rgb" if has_rgb else "x y z"
    k = 4 if has_rgb else 3
    rec = np.zeros(len(pts), dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4")] + ([("rgb", "<u4")] if has_rgb else [])))
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if has_rgb:
        c = np.clip(np.round(np.asarray(pointcloud.colors) * 255.0), 0, 255).astype(np.uint32)
        rec["rgb"] = (c[:, 0] << 16) | (c[:, 1] << 8) | c[:, 2]
    head = (f"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS {fields}\nSIZE {' '.join(['4'] * k)}\n"
            f"TYPE {'F F F U' if has_rgb else 'F F F'}\nCOUNT {' '.join(['1'] * k)}\nWIDTH {len(pts)}\nHEIGHT 1\n"
            f"VIEWPOINT 0 0 0 1 0 0 0\nPOINTS {len(pts)}\nDATA {'ascii' if write_ascii else 'binary'}\n")
    with open(filename, "wb") as f:
        f.write(head.encode("ascii"))
        if write_ascii:
            for r in rec:
                f.write((" ".join(repr(float(r[n])) if n != "rgb" else str(int(r[n])) for n in rec.dtype.names) + "\n").encode("ascii"))
        else:
            f.write(rec.tobytes())
    return True
