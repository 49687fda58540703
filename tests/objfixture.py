"""Writes a reference state from the golden file as an OBJ with UVs (v / vt / f a/b ...), at %.17g, so the
reference host program (and the same program with the CUDA plugins dropped in) can be started from it
through its own `input UV` path (main.cpp:1312-1377)."""
import numpy as np


def write_obj(path, V_rest, F, UV):
    V_rest, F, UV = np.asarray(V_rest), np.asarray(F), np.asarray(UV)
    uniq, inv = np.unique(V_rest, axis=0, return_inverse=True)         # weld the seam duplicates
    # keep first-occurrence order so the welded numbering is deterministic and close to the original
    first = np.full(len(uniq), len(V_rest), dtype=np.int64)
    np.minimum.at(first, inv, np.arange(len(V_rest)))
    order = np.argsort(first)
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[order] = np.arange(len(uniq))
    wid = rank[inv]
    with open(path, "w") as f:
        for p in uniq[order]:
            f.write("v %.17g %.17g %.17g\n" % tuple(p))
        for t in UV:
            f.write("vt %.17g %.17g\n" % tuple(t))
        for t in F:
            f.write("f %d/%d %d/%d %d/%d\n" % (wid[t[0]] + 1, t[0] + 1, wid[t[1]] + 1, t[1] + 1, wid[t[2]] + 1, t[2] + 1))
    return len(uniq)
