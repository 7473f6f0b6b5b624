"""Numpy prototype of the per-warp combine plan for the P1 matrix scatter ("next" item 1).  CPU only.

The P1 tet matrix kernel is bound by the rate of scalar fp64 REDs (profiles/assembly_r1.txt), so the lever is fewer REDs per tet.
Plan, built once per mesh in the symbolic phase for every chunk of 32 consecutive cells (one warp):
    rank[k]   uint16, k = lane * E + e (E = (d+1)^2 local entries): position of contribution k when the chunk's contributions are
              sorted by their destination slot in `vals`
    head[p]   1 bit per sorted position: first contribution of a destination (plus a forced head at every lane boundary p % E == 0,
              so each lane reduces a self-contained range of E sorted positions)
    dest[s]   uint32 per head: the destination slot
Numeric phase per warp: every lane forms its E local entries, writes them to shared memory at rank[k], then walks its own E sorted
positions adding up runs between heads and issues ONE RED per run.

This script emulates exactly that with numpy on a dolfin-layout cube, checks that the assembled values equal the plain scatter, and
counts the REDs and the plan bytes.   python tools/asm_plan_prototype.py [N] [mass]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fem_oracle as fo  # noqa: E402

WARP = 32


def build_plan(cells, row_ptr, col_idx):
    nc, nl = cells.shape
    E = nl * nl
    nchunk = (nc + WARP - 1) // WARP
    # destination slot of every contribution: row_ptr[row] + position of col in the row
    rows = np.repeat(cells, nl, axis=1).astype(np.int64)
    cols = np.tile(cells, (1, nl)).astype(np.int64)
    slot = np.empty((nc, E), dtype=np.int64)
    for c in range(nc):                       # prototype: clarity over speed
        for e in range(E):
            r = rows[c, e]
            seg = col_idx[row_ptr[r]:row_ptr[r + 1]]
            slot[c, e] = row_ptr[r] + np.searchsorted(seg, cols[c, e])
    pad = nchunk * WARP - nc
    slot_p = np.vstack([slot, np.full((pad, E), -1, dtype=np.int64)]) if pad else slot
    key = slot_p.reshape(nchunk, WARP * E)
    order = np.argsort(key, axis=1, kind="stable")                      # sorted position -> contribution
    rank = np.empty_like(order)
    np.put_along_axis(rank, order, np.arange(WARP * E)[None, :].repeat(nchunk, 0), axis=1)   # contribution -> sorted position
    skey = np.take_along_axis(key, order, axis=1)
    head = np.ones_like(skey, dtype=bool)
    head[:, 1:] = skey[:, 1:] != skey[:, :-1]
    head[:, ::E] = True                                                 # forced head at every lane boundary
    return {"rank": rank.astype(np.uint16), "head": head, "dest": skey, "E": E, "nchunk": nchunk, "slot": slot}


def assemble_with_plan(plan, Ke, nnz, skip_zero=True):
    """Emulates the numeric phase; returns (vals, number of REDs issued)."""
    nchunk, E = plan["nchunk"], plan["E"]
    nc = Ke.shape[0]
    vals = np.zeros(nnz)
    reds = 0
    flat = np.zeros((nchunk * WARP, E))
    flat[:nc] = Ke.reshape(nc, E)
    flat = flat.reshape(nchunk, WARP * E)
    for ch in range(nchunk):
        smem = np.zeros(WARP * E)
        smem[plan["rank"][ch]] = flat[ch]                                # scatter to sorted positions (shared memory)
        head, dest = plan["head"][ch], plan["dest"][ch]
        starts = np.nonzero(head)[0]
        sums = np.add.reduceat(smem, starts)                             # each lane walks its own E positions; runs end at heads
        d = dest[starts]
        live = (d >= 0) & ((sums != 0.0) if skip_zero else True)
        np.add.at(vals, d[live], sums[live])                             # one RED per run
        reds += int(live.sum())
    return vals, reds


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    mass = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    c, t = fo.unit_cube_mesh(N, N, N)
    nv, nc = c.shape[0], t.shape[0]
    rp, ci = fo.csr_pattern(t, nv)
    Ke = fo.local_laplace(c, t, 20.0) + (fo.local_mass(c, t, mass) if mass else 0.0)
    ref = fo.conform(fo.assemble_matrix(t, Ke, nv), rp, ci).data
    plan = build_plan(t, rp, ci)
    vals, reds = assemble_with_plan(plan, Ke, ci.size)
    err = np.abs(vals - ref).max() / np.abs(ref).max()
    plain = int((Ke != 0.0).sum())
    heads = int(plan["head"].sum())
    bytes_per_cell = (2 * WARP * plan["E"] + WARP * plan["E"] / 8 + 4 * heads / plan["nchunk"]) / WARP
    print("N=%d mass=%g: %d tets, plain scatter %d REDs (%.2f per tet), with the plan %d REDs (%.2f per tet, x%.2f fewer); "
          "max rel diff of the assembled values %.1e; plan %.0f B per tet (rank 2 B + head 1 bit per contribution, dest 4 B per run)"
          % (N, mass, nc, plain, plain / nc, reds, reds / nc, plain / max(reds, 1), err, bytes_per_cell))
    assert err < 1e-13


if __name__ == "__main__":
    main()
