"""Block schedule of the launches with the halo synchronisation folded in (lattice_qcd_rs_b200/csrc/lq_tuned.cuh:
lq_fold_geom / lq_fold_block; opt-in LQ_FLAG_FOLD_HALO_SYNC).  The device code cannot run here, so this is a line-by-line
Python restatement of its index arithmetic, checked over many geometries: every grid position maps to a distinct natural
block (a permutation), the blocks flagged `isb` are exactly those of the boundary (x2, x3) columns, and their number is
the count the last boundary block waits for before it releases the epoch (a mismatch would stall the neighbours)."""
import itertools


def geom(e, ghost, bs, zfirst):
    """lq_fold_geom: schedule of a kernel whose blocks hold `bs` consecutive sites; None: geometry not covered."""
    e0, e1, e2, e3 = e
    g2, g3 = ghost
    if not g2 and not g3:
        return None
    colsites = e0 * e1
    if colsites % bs or (colsites * e2 * e3) % bs:
        return None
    n3 = (2 if e3 >= 2 else 1) if g3 else 0  # boundary values of x3, x2
    n2 = (2 if e2 >= 2 else 1) if g2 else 0
    col_b = n3 * e2 + (e3 - n3) * n2
    bpc = colsites // bs
    total, n_b = e2 * e3 * bpc, col_b * bpc
    if n_b <= 0 or n_b > total:
        return None
    zf = 1 if (zfirst or n3 == 0) else 0
    n_f = n_b if zf else n3 * e2 * bpc
    sp = total // n_f
    return dict(bpc=bpc, nB=n_b, nF=n_f, S=min(max(sp, 1), 4), zfirst=zf, total=total, n2=n2, n3=n3)


def block(e, f, blk):
    """lq_fold_block: grid position -> (natural block index, block of a boundary column?)."""
    _, _, e2, e3 = e
    n2, n3, s, bpc = f["n2"], f["n3"], f["S"], f["bpc"]
    j = blk // s
    if blk - j * s == 0 and j < f["nF"]:
        isb = True
        cb = j // bpc
        within = j - cb * bpc
        p1 = n3 * e2
        if cb < p1:
            q = cb // e2
            x3 = 0 if q == 0 else e3 - 1
            x2 = cb - q * e2
        else:
            r = cb - p1
            q = r // n2
            x3 = (1 if n3 else 0) + q
            x2 = 0 if r - q * n2 == 0 else e2 - 1
        col = x2 + e2 * x3
    else:
        ji = blk - min((blk + s - 1) // s, f["nF"])
        ci = ji // bpc
        within = ji - ci * bpc
        if f["zfirst"]:
            isb = False
            i2 = e2 - n2
            q = ci // i2
            col = (1 if n2 else 0) + (ci - q * i2) + e2 * ((1 if n3 else 0) + q)
        else:
            q = ci // e2
            x2 = ci - q * e2
            isb = bool(n2 and x2 in (0, e2 - 1))
            col = x2 + e2 * ((1 if n3 else 0) + q)
    return col * bpc + within, isb


def test_fold_schedule_is_a_permutation_with_the_right_boundary_set():
    checked = 0
    for e2, e3 in itertools.product([1, 2, 3, 4, 8, 32], [1, 2, 3, 4, 8, 16, 32]):
        for ghost in [(0, 1), (1, 1), (1, 0)]:
            for (e0, e1, bs) in [(32, 32, 32), (32, 32, 128), (16, 8, 32), (16, 8, 128), (48, 48, 32), (48, 48, 128),
                                 (32, 4, 128), (8, 8, 32)]:
                for zf in (0, 1):
                    e = (e0, e1, e2, e3)
                    f = geom(e, ghost, bs, zf)
                    if f is None:
                        continue
                    checked += 1
                    seen, nb = set(), 0
                    for blk in range(f["total"]):
                        nat, isb = block(e, f, blk)
                        col = nat // f["bpc"]
                        x2, x3 = col % e2, col // e2
                        true_b = bool((ghost[0] and x2 in (0, e2 - 1)) or (ghost[1] and x3 in (0, e3 - 1)))
                        assert 0 <= nat < f["total"] and nat not in seen and true_b == bool(isb), (e, ghost, bs, zf, blk)
                        seen.add(nat)
                        nb += isb
                    assert len(seen) == f["total"] and nb == f["nB"], (e, ghost, bs, zf)
    assert checked > 1000
