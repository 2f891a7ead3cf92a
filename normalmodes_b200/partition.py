"""Node partitioning for multi-GPU runs.

The reference partitions the node graph with ParMETIS 4.0.3 (src/mod_geometry.f90:694-910), which is not
available here and cannot be reproduced bit-exactly without it; a partition is therefore an INPUT of this
package (`part[]`, e.g. read from the `*_vlist.dat` dump of a real reference run, src/mod_pevsl.f90:234-240).
For runs that start from raw mesh files this module supplies recursive coordinate bisection: NVSwitch gives
every GPU the same bandwidth to every peer, so only the halo volume and the balance matter, not the
neighbour topology."""
import numpy as np


def node_coordinates(mesh, fem):
    """Coordinates of all nn nodes (P2: edge nodes at edge midpoints) from the element->node table."""
    from .meshgen import element_nodes
    t2n = fem.t2n()
    Xn = element_nodes(mesh, fem.porder).reshape(-1, 3)
    X = np.empty((fem.nn, 3))
    X[t2n.ravel()] = Xn
    return X


def rcb(X, nparts, weights=None):
    """Recursive coordinate bisection into nparts (any positive integer) parts of ~equal weight."""
    n = X.shape[0]
    part = np.zeros(n, dtype=np.int32)
    w = np.ones(n) if weights is None else np.asarray(weights, dtype=float)

    def split(idx, p0, k):
        if k == 1:
            part[idx] = p0
            return
        kl = k // 2
        ext = X[idx].max(axis=0) - X[idx].min(axis=0)
        ax = int(np.argmax(ext))
        o = idx[np.argsort(X[idx, ax], kind="stable")]
        cw = np.cumsum(w[o])
        cut = int(np.searchsorted(cw, cw[-1] * kl / k))
        split(o[:cut], p0, kl)
        split(o[cut:], p0 + kl, k - kl)
    split(np.arange(n), 0, nparts)
    return part


def read_vlist(path, nn, vtxdist):
    """part[] from a reference `*_vlist.dat` (int32 original 1-based node id of each row block, ranks concatenated)."""
    vl = np.fromfile(path, dtype="<i4")
    assert vl.size == nn
    part = np.empty(nn, dtype=np.int32)
    for r in range(len(vtxdist) - 1):
        part[vl[vtxdist[r]:vtxdist[r + 1]] - 1] = r
    return part
