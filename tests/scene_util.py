"""Shared helpers for the parity tests (test infrastructure)."""
import gzip
import os

import numpy as np

from oracle import pyoracle as po
from scenes import library

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_ppm(name):
    with gzip.open(os.path.join(GOLDEN, name + ".ppm.gz"), "rb") as f:
        return po.read_ppm(f.read())


def oracle_scene(desc, flags=None):
    return library.apply(desc, po.OracleScene(flags=desc.flags if flags is None else flags))


def ray_key(r):
    """Order-independent key of a ray record: the 64 payload bytes (queue order is nondeterministic in the
    reference, SURVEY Q20, so queues are compared as sorted multisets)."""
    raw = np.ascontiguousarray(r).view(np.uint8).reshape(len(r), 80)[:, :64]
    return raw


def sorted_payload(r, ignore_depth_of_shadow=False):
    raw = ray_key(r).copy()
    if len(raw) == 0:
        return raw
    order = np.lexsort(raw.T[::-1])
    return raw[order]


def random_rays(bbox, n, seed=1, spread=1.5):
    rng = np.random.default_rng(seed)
    rr = np.zeros(n, po.RAY_DTYPE)
    lo, hi = np.asarray(bbox[:3], np.float64), np.asarray(bbox[3:], np.float64)
    ext = hi - lo
    rr["origin"] = (lo - (spread - 1) / 2 * ext + ext * spread * rng.random((n, 3))).astype(np.float32)
    dd = rng.normal(size=(n, 3))
    dd /= np.linalg.norm(dd, axis=1)[:, None]
    rr["direction"] = dd.astype(np.float32)
    rr["t_min"], rr["t_max"], rr["t"], rr["w"], rr["depth"] = 1e-6, np.finfo(np.float32).max, np.finfo(np.float32).max, 1.0, 1
    rr["id"] = np.arange(n)
    return rr


def grazing_rays(verts, faces, reps=2000, seed=7):
    """Rays launched from 1e-6 .. 1e-4 above / below random points of every face, towards it (hit t just above
    RAY_EPSILON): the case in which a non-conservative child box loses a hit."""
    rng = np.random.default_rng(seed)
    v = np.asarray(verts, np.float32).reshape(-1, 3)
    tri = v[np.asarray(faces, np.int64)]
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]).astype(np.float64)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
    w = rng.dirichlet((1, 1, 1), size=(len(tri), reps))
    p = np.einsum("frk,fkc->frc", w, tri.astype(np.float64))
    eps = 10.0 ** rng.uniform(-6, -4, size=(len(tri), reps, 1))
    side = rng.choice([-1.0, 1.0], size=(len(tri), reps, 1))
    org = p + side * eps * n[:, None, :]
    dirs = -side * n[:, None, :] + rng.normal(0, 0.3, size=org.shape)
    dirs /= np.linalg.norm(dirs, axis=2, keepdims=True)
    rr = np.zeros(len(tri) * reps, po.RAY_DTYPE)
    rr["origin"] = org.reshape(-1, 3).astype(np.float32)
    rr["direction"] = dirs.reshape(-1, 3).astype(np.float32)
    rr["t_min"], rr["t_max"], rr["t"], rr["w"], rr["depth"] = 1e-6, np.finfo(np.float32).max, np.finfo(np.float32).max, 1.0, 1
    rr["id"] = np.arange(len(rr))
    return rr
