"""numpy model of the multigrid preconditioner of the PPE solve (openmps_b200/csrc/mps_mg.cu, mps_cg.cu k_pcg_stream), used by
tests/test_multigrid.py: the hierarchy built on the host from the matrix and the cell keys, the V-cycle written out with the
device's tables, and a textbook PCG with the reference's stopping rule (Computer.hpp:1382-1428).  Test infrastructure."""
import numpy as np
import scipy.sparse as sp

NONE = 0xFFFFFFFF


def stencil_offsets(dim):
    rng = (-1, 0, 1)
    if dim == 2:
        return [(ox, oz) for ox in rng for oz in rng]
    return [(ox, oy, oz) for ox in rng for oy in rng for oz in rng]


def encode(coords, dims):
    key = np.zeros(len(coords), np.int64)
    for a in range(coords.shape[1]):
        key = key * int(dims[a]) + coords[:, a]
    return key


def decode(key, dims):
    key = np.asarray(key, np.int64).copy()
    out = np.zeros((len(key), len(dims)), np.int64)
    for a in range(len(dims) - 1, -1, -1):
        out[:, a] = key % int(dims[a]); key //= int(dims[a])
    return out


def host_hierarchy(A_slot, row_key, dims0, levels, omega):
    """The hierarchy as mps_mg.cu defines it, from the slot-ordered matrix and every row's cell key (-1: Disabled).
    Returns a list of dicts per level: key, nbr, child, parent, S, dinv."""
    dim = len(dims0)
    K = 3 ** dim
    offs = np.array(stencil_offsets(dim), np.int64)
    out = []
    in_cell = row_key >= 0
    keys0 = np.unique(row_key[in_cell])
    cell_of_row = np.full(len(row_key), -1, np.int64)
    cell_of_row[in_cell] = np.searchsorted(keys0, row_key[in_cell])
    n0 = len(keys0)
    rows = np.flatnonzero(in_cell)
    P = sp.csr_matrix((np.ones(len(rows)), (rows, cell_of_row[rows])), shape=(len(row_key), n0))
    Ac = (P.T @ A_slot @ P).tocsr()
    keys, dims = keys0, np.array(dims0, np.int64)
    for l in range(levels):
        n = len(keys)
        coords = decode(keys, dims)
        S = np.zeros((n, K)); nbr = np.full((n, K), NONE, np.uint32)
        Acoo = Ac.tocoo()
        look = {int(k): i for i, k in enumerate(keys)}
        for s, o in enumerate(offs):
            nb = coords + o
            ok = np.all((nb >= 0) & (nb < dims), axis=1)
            nk = encode(np.where(ok[:, None], nb, 0), dims)
            ids = np.array([look.get(int(k), -1) if f else -1 for k, f in zip(nk, ok)], np.int64)
            nbr[ids >= 0, s] = ids[ids >= 0]
        # scatter the Galerkin entries into stencil slots
        dcoord = coords[Acoo.col] - coords[Acoo.row]
        assert np.abs(dcoord).max(initial=0) <= 1, "coarse operator is not a 3^D stencil"
        slot = np.zeros(len(Acoo.data), np.int64)
        for a in range(dim):
            slot = slot * 3 + (dcoord[:, a] + 1)
        np.add.at(S, (Acoo.row, slot), Acoo.data)
        centre = S[:, K // 2]
        dinv = np.where(centre != 0, omega / np.where(centre != 0, centre, 1.0), 0.0)
        lv = {"key": keys.astype(np.uint32), "nbr": nbr, "S": S, "dinv": dinv, "dims": dims.copy()}
        out.append(lv)
        if l + 1 == levels:
            break
        pdims = (dims + 1) // 2
        pk = encode(coords >> 1, pdims)
        pkeys = np.unique(pk)
        parent = np.searchsorted(pkeys, pk)
        lv["parent"] = parent.astype(np.uint32)
        Pl = sp.csr_matrix((np.ones(n), (np.arange(n), parent)), shape=(n, len(pkeys)))
        Ac = (Pl.T @ Ac @ Pl).tocsr()
        keys, dims = pkeys, pdims
    return out, cell_of_row


def _apply(S, nbr, e):
    pad = np.append(e, 0.0)
    idx = np.where(nbr == NONE, len(e), nbr).astype(np.int64)
    return (S * pad[idx]).sum(1)


def vcycle(levels, r0, gamma, top_sweeps, top_cells):
    """One V(1,1) cycle exactly as mg_vcycle (mps_cg.cu) runs it; levels[l] needs S, nbr, dinv, parent."""
    L = 1
    while L < len(levels) and len(levels[L - 1]["dinv"]) > top_cells:
        L += 1
    r = [None] * L; e = [None] * L
    r[0] = r0; e[0] = levels[0]["dinv"] * r0
    for l in range(L - 1):
        lo = levels[l]
        res = r[l] - _apply(lo["S"], lo["nbr"], e[l])
        r[l + 1] = np.bincount(lo["parent"].astype(np.int64), weights=res, minlength=len(levels[l + 1]["dinv"]))
        e[l + 1] = levels[l + 1]["dinv"] * r[l + 1]
    top = levels[L - 1]
    for _ in range(top_sweeps):
        e[L - 1] = e[L - 1] + top["dinv"] * (r[L - 1] - _apply(top["S"], top["nbr"], e[L - 1]))
    for l in range(L - 2, -1, -1):
        lv = levels[l]
        et = e[l] + gamma * e[l + 1][lv["parent"].astype(np.int64)]
        e[l] = et + lv["dinv"] * (r[l] - _apply(lv["S"], lv["nbr"], et))
    return e[0], L


def pcg(A, b, x0, eps, M, maxit=None):
    x = x0.copy()
    r = b - A @ x
    tol = (r @ r) * eps * eps
    if tol == 0:
        return x, 0
    z = M(r); p = z.copy(); rz = r @ z
    for it in range(1, (maxit or len(b)) + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p; r -= alpha * Ap
        if r @ r < tol:
            return x, it
        z = M(r); rz_new = r @ z
        p = z + (rz_new / rz) * p; rz = rz_new
    return x, maxit or len(b)
