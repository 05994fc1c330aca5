"""z-slab partition of the cylindrical mesh for the multi-GPU solve of ONE large system (BASELINE config 5).

The reference has no counterpart (single process; run.py:50-53).  k = z is the slowest index of every vector block
(utils.py:41-47), so the part of a block that belongs to a range of z levels is one contiguous run.  Rank r owns
the global cell levels [zsplit[r], zsplit[r+1]) and the node levels of the same range (the last rank also the top
node level); its GPU context works on the local mesh of the cells [max(z_lo - 1, 0), z_hi): one ghost cell layer
below, one ghost node level on each side (csrc/cs_dist.cu).  These helpers are the host-side bookkeeping: the split,
global <-> local vector layout, owned parts, and which rank factors which theta mode.
"""
import numpy as np

from .mesh import CylMesh


def split_levels(nz, nranks, min_levels=2):
    """balanced contiguous split of the nz cell levels: boundaries zsplit[0..nranks]"""
    if nz < min_levels * nranks:
        raise ValueError("{} z levels cannot give {} ranks at least {} levels each".format(nz, nranks, min_levels))
    return [int(round(r * nz / float(nranks))) for r in range(nranks + 1)]


def local_range(zsplit, rank):
    """global cell levels [k0, k1) covered by the local mesh of `rank`"""
    return max(zsplit[rank] - 1, 0), zsplit[rank + 1]


def local_mesh(mesh, k0, k1):
    zn = mesh.vectorNz
    return CylMesh([mesh.hx, mesh.hy, mesh.hz[k0:k1]], x0=[0.0, 0.0, zn[k0]])


def blocks(mesh, family):
    """(entries per level, node-level?) of the component blocks of a cell (0) / edge (1) / face (2) vector"""
    nxy = int(mesh.vnC[0] * mesh.vnC[1])
    if family == 0:
        return [(nxy, False)]
    if family == 1:
        return [(nxy, True)] if mesh.isSymmetric else [(nxy, True), (nxy, True), (nxy + 1, False)]
    return [(nxy, False), (nxy, True)] if mesh.isSymmetric else [(nxy, False), (nxy, False), (nxy, True)]


def _split_blocks(vec, mesh, family, nz):
    out, o = [], 0
    for per, node in blocks(mesh, family):
        nlev = nz + 1 if node else nz
        out.append(vec[o:o + per * nlev].reshape(nlev, per))
        o += per * nlev
    assert o == vec.shape[0], "vector does not match the mesh / family"
    return out


def to_local(vec, mesh, family, k0, k1):
    """the part of a GLOBAL vector that lives on the local mesh [k0, k1), in the local layout"""
    vec = np.asarray(vec)
    parts = []
    for (per, node), blk in zip(blocks(mesh, family), _split_blocks(vec, mesh, family, int(mesh.vnC[2]))):
        parts.append(blk[k0:k1 + (1 if node else 0)].reshape(-1))
    return np.concatenate(parts)


def to_local_sparse(idx, val, mesh, family, k0, k1):
    """to_local for a sparse global vector (indices idx, values val; e.g. a source s_e with a handful of wire faces): the
    dense LOCAL vector, without ever forming the global one (config 5: 142 M faces)"""
    idx, val = np.asarray(idx, dtype=np.int64), np.asarray(val)
    nz = int(mesh.vnC[2])
    out, g_off, l_off = [], 0, 0
    pieces = []
    for per, node in blocks(mesh, family):
        nlev_g = nz + 1 if node else nz
        lo, hi = k0, k1 + (1 if node else 0)
        sel = (idx >= g_off + lo * per) & (idx < g_off + hi * per)
        pieces.append((l_off + idx[sel] - g_off - lo * per, val[sel]))
        g_off += per * nlev_g
        l_off += per * (hi - lo)
    loc = np.zeros(l_off, dtype=val.dtype if val.size else float)
    for li, lv in pieces:
        loc[li] = lv
    return loc


def owned_part(local_vec, mesh, family, zsplit, rank):
    """per block, the owned levels of a LOCAL vector (what a gather collects)"""
    k0, k1 = local_range(zsplit, rank)
    z_lo, z_hi, nz = zsplit[rank], zsplit[rank + 1], int(mesh.vnC[2])
    out = []
    lm_nz = k1 - k0
    for (per, node), blk in zip(blocks(mesh, family), _split_blocks(np.asarray(local_vec), mesh, family, lm_nz)):
        hi = z_hi + (1 if (node and z_hi == nz) else 0)
        out.append(blk[z_lo - k0:hi - k0])
    return out


def assemble(owned_parts, mesh, family):
    """global vector from the owned parts of all ranks (list over ranks of owned_part results)"""
    cols = []
    for b in range(len(blocks(mesh, family))):
        cols.append(np.concatenate([p[b] for p in owned_parts], axis=0).reshape(-1))
    return np.concatenate(cols)


def mode_owner(nth, nranks):
    """owner rank and position (among the owner's mode vectors) of every theta mode vector m = 0 .. nth-1 of a complex
    field: mode m is solved with the stored factor h = min(m, nth - m), factor h lives on rank h % nranks"""
    owner, pos, cnt = [], [], [0] * nranks
    for m in range(nth):
        h = 0 if nth == 1 else min(m, nth - m)
        o = h % nranks
        owner.append(o)
        pos.append(cnt[o])
        cnt[o] += 1
    return owner, pos
