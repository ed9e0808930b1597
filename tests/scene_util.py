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
