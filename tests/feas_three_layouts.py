#!/usr/bin/env python
"""tests/feas_three_layouts.py -- host-side measurement for FeAs (config 4) on several GPUs without the full-vector gather.

The Hubbard two-layout exchange works because every off-diagonal term changes ONE of the two words of a
row: down hops stay inside a row shard, up hops inside a column shard.  FeAsBasedSc adds on-site two-spin terms
(FeBasedSc.h:376-432: the spin flip  c+_{g,up} c_{g,dn} c+_{g',dn} c_{g',up}  and the pair hop  c+_{g,up} c+_{g,dn} c_{g',dn} c_{g',up})
that change BOTH words.  They flip the same two bits in both, so  w = up XOR down  is invariant under them: a third layout sharded
by w keeps them local.  This script takes the oracle's CRS matrix (oracle/lanczos_oracle.c, pinned to the reference's own
FeBasedSc.h by tests/test_reference_pin.py), sorts every off-diagonal entry into

    down hop  (up word unchanged)      -> local in the COLUMN layout (a range of up states with all their down states)
    up hop    (down word unchanged)    -> local in the ROW layout    (a range of down states with all their up states)
    two-spin  (both words change)      -> local in the XOR layout, if  up XOR down  is the same on both sides

and reports, for N ranks, the vector elements a rank needs from other ranks under (a) the row-sharded gather scheme of today
(distinct source rows outside the own shard) and (b) the three-layout scheme (four all-to-alls of the local vector per mat-vec).
Used by tests/test_oracle.py::test_feas_two_spin_terms_conserve_up_xor_down.

    python tests/feas_three_layouts.py [lx ly nup ndown [nranks]]      (default: 2 x 3 lattice, 4 up 4 down, 8 ranks)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc          # noqa: E402
from tests import cases                   # noqa: E402


def classify(o):
    """-> dict with the entry classes of the oracle's matrix; every array is indexed by off-diagonal entry."""
    rowptr, colind, vals = o.crs()
    n = o.rows()
    up, dn = o.row_words(0), o.row_words(1)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    off = rows != colind
    r, c, v = rows[off], colind[off], vals[off]
    dup, ddn = up[r] ^ up[c], dn[r] ^ dn[c]
    kind = np.where((dup == 0) & (ddn != 0), 0, np.where((dup != 0) & (ddn == 0), 1, 2))     # 0 down hop, 1 up hop, 2 two-spin
    return dict(n=n, rows=r, cols=c, vals=v, kind=kind, dup=dup, ddn=ddn, up=up, dn=dn, rowptr=rowptr, colind=colind, allvals=vals)


def halo_of_gather(cl, nranks):
    """distinct off-shard source rows per rank when the rows are split into nranks contiguous shards (today's scheme for FeAs)"""
    n = cl["n"]
    bounds = [(n * k) // nranks for k in range(nranks + 1)]
    out = []
    for k in range(nranks):
        sel = (cl["rows"] >= bounds[k]) & (cl["rows"] < bounds[k + 1])
        src = np.unique(cl["cols"][sel])
        out.append(int(((src < bounds[k]) | (src >= bounds[k + 1])).sum()))
    return out, bounds


def owners(cl, nup_states, nranks):
    """owner rank of every vector element in the three layouts.  idx = iup + idn * Nup (BasisFeAsBasedSc perfectIndex order)."""
    n = cl["n"]
    idx = np.arange(n, dtype=np.int64)
    iup, idn = idx % nup_states, idx // nup_states
    ndn_states = n // nup_states
    row_owner = (idn * nranks) // ndn_states                      # ROW layout: ranges of down states, every up state of them
    col_owner = (iup * nranks) // nup_states                      # COLUMN layout: ranges of up states, every down state of them
    w = cl["up"] ^ cl["dn"]
    classes, inv, counts = np.unique(w, return_inverse=True, return_counts=True)
    load = np.zeros(nranks, dtype=np.int64)                       # XOR layout: whole classes, largest first onto the lightest rank
    cls_owner = np.zeros(len(classes), dtype=np.int64)
    for c in np.argsort(-counts, kind="stable"):
        r = int(np.argmin(load))
        cls_owner[c] = r
        load[r] += counts[c]
    return row_owner, col_owner, cls_owner[inv], load


def three_layout_matvec(cl, nup_states, nranks, y):
    """x = H y the way N ranks would do it: every rank applies a class of entries only to the elements it owns in the layout that
    keeps the class local, and reads only elements it owns in that layout.  Returns (x, elements moved per all-to-all)."""
    n = cl["n"]
    row_owner, col_owner, xor_owner, _ = owners(cl, nup_states, nranks)
    layout_of_kind = (col_owner, row_owner, xor_owner)            # down hops, up hops, two-spin terms
    x = np.zeros(n)
    rows_all = np.repeat(np.arange(n), np.diff(cl["rowptr"]))
    diag = rows_all == cl["colind"]
    np.add.at(x, rows_all[diag], cl["allvals"][diag] * y[cl["colind"][diag]])     # the diagonal is local in every layout
    for k in range(3):
        own = layout_of_kind[k]
        sel = cl["kind"] == k
        r, c = cl["rows"][sel], cl["cols"][sel]
        if not np.array_equal(own[r], own[c]):
            raise AssertionError("an entry of class %d joins two ranks in its layout" % k)
        np.add.at(x, r, cl["vals"][sel] * y[c])
    # an all-to-all between the row layout and another one moves the elements whose owner differs
    moved = [int((row_owner != col_owner).sum()), int((row_owner != xor_owner).sum())]
    return x, moved


def main(argv):
    lx, ly, nup, ndn = (int(a) for a in argv[1:5]) if len(argv) >= 5 else (2, 3, 4, 4)
    nranks = int(argv[5]) if len(argv) >= 6 else 8
    case = cases.feas_cluster(lx, ly, nup, ndn)
    o = cases.make_oracle(orc, case, fast_rank=1)
    cl = classify(o)
    n = cl["n"]
    names = ("down hops (column layout)", "up hops (row layout)", "two-spin terms (XOR layout)")
    print("FeAs %dx%d, 2 orbitals, %d up %d down: %d rows, %d off-diagonal entries" % (lx, ly, nup, ndn, n, len(cl["kind"])))
    for k in range(3):
        print("  %-30s %9d entries (%.2f per row)" % (names[k], int((cl["kind"] == k).sum()), (cl["kind"] == k).sum() / n))
    two = cl["kind"] == 2
    same = (cl["up"][cl["rows"][two]] ^ cl["dn"][cl["rows"][two]]) == (cl["up"][cl["cols"][two]] ^ cl["dn"][cl["cols"][two]])
    print("  two-spin entries with  up XOR down  equal on both sides: %d of %d" % (int(same.sum()), int(two.sum())))
    print("  two-spin entries that flip the same bits in both words:   %d of %d" % (int((cl["dup"][two] == cl["ddn"][two]).sum()), int(two.sum())))
    halo, bounds = halo_of_gather(cl, nranks)
    loc = [bounds[k + 1] - bounds[k] for k in range(nranks)]
    print("  %d ranks, row shards of about %d rows:" % (nranks, n // nranks))
    print("    gather scheme of today: a rank needs %d..%d distinct remote rows = %.2f x its own shard (it fetches all %d)"
          % (min(halo), max(halo), float(np.mean(halo)) / float(np.mean(loc)), n - min(loc)))
    print("    three layouts: 4 all-to-alls (row->column, column->row, row->XOR, XOR->row) of the %d local rows each = %.2f x its own shard"
          % (int(np.mean(loc)), 4.0 * (nranks - 1) / nranks))
    nup_states = len(o.basis(0))
    y = np.cos(0.37 * np.arange(n) + 0.1)
    x, moved = three_layout_matvec(cl, nup_states, nranks, y)
    xref = np.zeros(n)
    o.matvec(xref, y, faithful=False)
    _, _, _, load = owners(cl, nup_states, nranks)
    print("    emulated on %d ranks: max |x - oracle| = %.2e; elements that change owner row<->column %d, row<->XOR %d of %d; XOR shards %d..%d rows"
          % (nranks, float(np.abs(x - xref).max()), moved[0], moved[1], n, int(load.min()), int(load.max())))
    classes = np.unique(cl["up"] ^ cl["dn"])
    sizes = np.array([int(((cl["up"] ^ cl["dn"]) == w).sum()) for w in classes])
    print("    XOR classes: %d, sizes %d..%d (largest = %.3f of a shard): whole classes can be dealt to the ranks"
          % (len(classes), sizes.min(), sizes.max(), sizes.max() / float(np.mean(loc))))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
