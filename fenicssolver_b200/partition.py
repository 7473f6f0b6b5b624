"""Node partition of an arbitrary simplex mesh for the distributed path (one process per GPU).

The reference distributes through dolfin's SCOTCH partitioner (SolverBase.py:102-118, 634); here the nodes of the
space (vertices for degree 1, vertices + edge nodes for degree 2) are split by recursive coordinate bisection, every
rank keeps the cells that touch one of its nodes (owner computes: the rows of owned nodes receive all their
contributions locally, so assembly needs no communication) and numbers its nodes PETSc-style:

    local ids  [0, n_owned)          owned nodes, ascending global id
               [n_owned, n_local)    ghost nodes, grouped by owner rank, ascending global id inside a group

so a neighbour's ghosts form one contiguous range that a receive fills in place, and the matching send list on the
owner's side is the same global ids in the same order.  Every rank derives all lists from the replicated host mesh
with the same deterministic code, so no handshake is needed.
"""
from __future__ import annotations

import numpy as np


def rcb_partition(coords, nparts):
    """Recursive coordinate bisection: part id per point.  Splits the longest axis of each group at the weighted
    median (sizes proportional to the number of parts on each side); ties are broken by point index, so the result is
    a pure function of the coordinates."""
    coords = np.asarray(coords, dtype=np.float64)
    part = np.zeros(coords.shape[0], dtype=np.int32)

    def split(idx, p0, np_):
        if np_ == 1:
            part[idx] = p0
            return
        left_parts = np_ // 2
        x = coords[idx]
        axis = int(np.argmax(x.max(axis=0) - x.min(axis=0)))
        order = np.lexsort((idx, x[:, axis]))
        nleft = int(round(idx.size * left_parts / np_))
        split(idx[order[:nleft]], p0, left_parts)
        split(idx[order[nleft:]], p0 + left_parts, np_ - left_parts)

    split(np.arange(coords.shape[0], dtype=np.int64), 0, int(nparts))
    return part


class NodePartition:
    """This rank's view of a node partition.  cell_nodes[ncells][nl] are global node ids; part[nnodes] the owner ranks."""

    def __init__(self, cell_nodes, part, rank, nranks):
        cell_nodes = np.asarray(cell_nodes, dtype=np.int64)
        part = np.asarray(part, dtype=np.int32)
        self.rank, self.nranks = int(rank), int(nranks)
        nnodes = part.size
        owner = part[cell_nodes]                                        # [nc, nl]
        # (rank on which the node is a ghost, node): a node is a ghost on every other rank owning a node of one of its cells
        nl = cell_nodes.shape[1]
        mixed = np.nonzero((owner != owner[:, :1]).any(axis=1))[0]
        om, cm = owner[mixed], cell_nodes[mixed]
        on_rank = np.repeat(om, nl, axis=1).ravel()                     # for each (a, b): owner of node a ...
        node = np.tile(cm, (1, nl)).ravel()                             # ... sees node b
        keep = on_rank != part[node]
        pairs = np.unique(on_rank[keep].astype(np.int64) * nnodes + node[keep])
        g_rank, g_node = pairs // nnodes, pairs % nnodes                # sorted by (ghost-on rank, global id)
        self.owned = np.nonzero(part == rank)[0].astype(np.int64)
        mine = g_rank == rank
        ghosts, gowner = g_node[mine], part[g_node[mine]]
        order = np.lexsort((ghosts, gowner))
        self.ghosts, gowner = ghosts[order], gowner[order]
        self.n_owned, self.n_local = self.owned.size, self.owned.size + self.ghosts.size
        self.l2g = np.concatenate([self.owned, self.ghosts])
        self.g2l = np.full(nnodes, -1, dtype=np.int64)
        self.g2l[self.l2g] = np.arange(self.n_local)
        # receive ranges
        nb_recv = np.unique(gowner)
        to_me = part[g_node] == rank                                    # my nodes that are ghosts elsewhere
        nb_send = np.unique(g_rank[to_me])
        self.neighbours = np.union1d(nb_recv, nb_send).astype(np.int32)
        self.recv_off = np.zeros(self.neighbours.size, dtype=np.int64)
        self.recv_cnt = np.zeros(self.neighbours.size, dtype=np.int64)
        send_lists = []
        for i, r in enumerate(self.neighbours):
            sel = np.nonzero(gowner == r)[0]
            self.recv_cnt[i] = sel.size
            self.recv_off[i] = self.n_owned + (sel[0] if sel.size else 0)
            send_lists.append(self.g2l[g_node[to_me & (g_rank == r)]])    # ascending global id = r's ghost order for me
        self.send_ptr = np.concatenate([[0], np.cumsum([s.size for s in send_lists])]).astype(np.int64)
        self.send_idx = np.concatenate(send_lists).astype(np.int64) if send_lists else np.zeros(0, dtype=np.int64)
        # local cells: every cell with an owned node, in global cell order, renumbered
        self.cells_global = np.nonzero((owner == rank).any(axis=1))[0]
        self.cell_nodes_local = self.g2l[cell_nodes[self.cells_global]]
        assert self.cell_nodes_local.min(initial=0) >= 0

    def to_local(self, gids):
        """Global node ids -> (local ids, mask of those present on this rank)."""
        g = np.asarray(gids, dtype=np.int64)
        l = self.g2l[g]
        return l, l >= 0
