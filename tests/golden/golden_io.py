"""Serialisation helpers of the golden vectors (tests/golden/*.json, *.npz).  Floats are stored as Python reprs in JSON,
which round-trip binary64 exactly."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
QUERIES = os.path.join(HERE, "random_queries.json")
CONFIG1 = os.path.join(HERE, "config1_r%d.npz")
CONFIG1_RADII = (50, 500, 2500)
QUERIES_PER_GRID = 200
SEED = 777


def query_to_dict(q):
    d = {}
    if q.SpotsAOI is not None:
        d["spots"] = [[s.X, s.Z] for s in q.SpotsAOI.Spots]
        d["spot_dists"] = [int(v) for v in q.SpotsAOI.Dists]
    if q.BoxAOI is not None:
        d["box"] = [q.BoxAOI.Center.X, q.BoxAOI.Center.Z, q.BoxAOI.Extent.X, q.BoxAOI.Extent.Z]
    if q.SphereAOI is not None:
        d["sphere"] = [q.SphereAOI.Center.X, q.SphereAOI.Center.Z, q.SphereAOI.Radius]
    if q.ConeAOI is not None:
        c = q.ConeAOI
        d["cone"] = [c.Center.X, c.Center.Z, c.Direction.X, c.Direction.Z, c.Angle, c.Radius]
    return d


def dict_to_query(d):
    from channeld_b200 import controller as C

    q = C.SpatialInterestQuery()
    if "spots" in d:
        q.SpotsAOI = C.SpotsAOI(Spots=[C.SpatialInfo(X=x, Z=z) for x, z in d["spots"]], Dists=list(d["spot_dists"]))
    if "box" in d:
        b = d["box"]
        q.BoxAOI = C.BoxAOI(Center=C.SpatialInfo(X=b[0], Z=b[1]), Extent=C.SpatialInfo(X=b[2], Z=b[3]))
    if "sphere" in d:
        s = d["sphere"]
        q.SphereAOI = C.SphereAOI(Center=C.SpatialInfo(X=s[0], Z=s[1]), Radius=s[2])
    if "cone" in d:
        c = d["cone"]
        q.ConeAOI = C.ConeAOI(Center=C.SpatialInfo(X=c[0], Z=c[1]), Direction=C.SpatialInfo(X=c[2], Z=c[3]), Angle=c[4], Radius=c[5])
    return q


def oracle_kwargs(d):
    kw = {}
    if "spots" in d:
        kw["spots"] = [tuple(p) for p in d["spots"]]
        kw["spot_dists"] = list(d["spot_dists"])
    for k in ("box", "sphere", "cone"):
        if k in d:
            kw[k] = tuple(d[k])
    return kw


def load_queries():
    with open(QUERIES) as f:
        return json.load(f)


def load_config1(radius):
    return dict(np.load(CONFIG1 % radius))
