"""Host-side schedule of the tcgen05 block-sparse GEMM (lcn_gemm_tc.cu: tc_fill_schedule): K-chunk order, operand
ring placement and accumulator completion order, checked by replaying the producer / MMA protocol on the CPU.
No GPU needed: lcn_debug_tc_schedule only runs the host code that fills the kernel parameters."""
import ctypes as C

import numpy as np
import pytest

from oracle import lcn_oracle as O
from lcn_pose_b200 import _lib as L

MAXC, GMAX, RING = 34, 6, 23


def _schedule(support, FC, tiles, sms=148):
    lib = L.load()
    kmask = (C.c_uint32 * 17)(*[int(sum(1 << j for j in range(17) if support[i, j])) for i in range(17)])
    sched = (C.c_uint32 * (MAXC * MAXC))()
    nit = (C.c_uint8 * MAXC)()
    goc0 = (C.c_uint8 * (MAXC + 1))()
    order = (C.c_uint8 * (MAXC * GMAX))()
    runs = (C.c_uint64 * (MAXC * MAXC))()
    dcnt = (C.c_uint8 * (MAXC * GMAX))()
    lib.lcn_debug_tc_schedule.restype = C.c_int
    ng = lib.lcn_debug_tc_schedule(kmask, FC, tiles, sms, sched, nit, goc0, order, runs, dcnt)
    _schedule.runs = np.array(runs[:], dtype=np.uint64).reshape(MAXC, MAXC)
    _schedule.dcnt = np.array(dcnt[:]).reshape(MAXC, GMAX)
    return ng, np.array(sched[:]).reshape(MAXC, MAXC), list(nit), list(goc0), np.array(order[:]).reshape(MAXC, GMAX)


@pytest.mark.parametrize("knn,FC,tiles", [(1, 1, 32), (2, 1, 32), (3, 1, 32), (3, 2, 128), (0, 1, 32), (0, 2, 8), (3, 1, 1)])
def test_schedule_protocol(knn, FC, tiles):
    sup = (O.get_neighbour_matrix_by_hand(knn=knn).T != 0) if knn > 0 else np.ones((17, 17), bool)
    ng, sched, nit, goc0, order = _schedule(sup, FC, tiles)
    NC = 17 * FC
    assert goc0[0] == 0 and goc0[ng] == NC
    total_blocks = 0
    for g in range(ng):
        oc0, G = goc0[g], goc0[g + 1] - goc0[g]
        assert 1 <= G <= GMAX
        present = {kc: [q for q in range(G) if sup[kc // FC, (oc0 + q) // FC]] for kc in range(NC)}
        want = {kc for kc, qs in present.items() if qs}
        seen, regions, last_touch, done_at, freed = [], [], {}, {}, -1
        for it in range(nit[g]):
            s = int(sched[g, it])
            kc, bits, off, wait, done = s & 63, (s >> 6) & 63, (s >> 12) & 31, ((s >> 17) & 127) - 1, (s >> 24) & 63
            assert bits == sum(1 << q for q in present[kc]), (g, it)
            need = 2 + bin(bits).count("1")
            assert off + need <= RING
            assert wait < it
            # producer protocol: the producer has waited for iterations <= freed (MMAs retire in order); every newer
            # earlier iteration may still be live and must not overlap
            freed = max(freed, wait)
            for j in range(freed + 1, it):
                o, n = regions[j]
                assert not (o < off + need and off < o + n), (g, it, j)
            # and nothing is waited for needlessly far: `wait` itself overlaps (or there is no wait)
            if wait >= 0:
                o, n = regions[wait]
                assert o < off + need and off < o + n
            # MMA runs: cover exactly the present chunks, adjacent, <= 4 long (accumulators start zeroed: all accumulate)
            rw = int(_schedule.runs[g, it])
            covered = 0
            for r in range(rw >> 60):
                e = (rw >> (6 * r)) & 63
                q, ln = e & 7, ((e >> 3) & 3) + 1
                for c in range(q, q + ln):
                    assert (bits >> c) & 1 and not (covered >> c) & 1
                    covered |= 1 << c
            assert covered == bits
            regions.append((off, need))
            seen.append(kc)
            total_blocks += bin(bits).count("1")
            for q in present[kc]:
                last_touch[(q, it % 2)] = it               # two MMA warps: even / odd iterations
            for q in range(G):
                if (done >> q) & 1:
                    assert (q, it % 2) not in done_at
                    done_at[(q, it % 2)] = it
        assert sorted(seen) == sorted(want) and len(set(seen)) == len(seen)
        assert done_at == last_touch                       # each warp commits done[q] right after its last MMA into q
        for q in range(G):                                 # and done[q] expects exactly those arrivals
            assert _schedule.dcnt[g, q] == sum((q, w) in last_touch for w in (0, 1))
        comp = [max(done_at.get((int(q), 0), -1), done_at.get((int(q), 1), -1)) for q in order[g, :G]]
        assert sorted(int(q) for q in order[g, :G]) == list(range(G))
        assert comp == sorted(comp)                        # the epilogue visits chunks in completion order
    assert total_blocks == int(sup.sum()) * FC * FC


def test_knn3_chunks_complete_staggered():
    """The point of the ordering: for the knn=3 mask at least one chunk of every group completes in the first 60 % of
    the group's block MMAs, so its epilogue overlaps the main loop."""
    sup = O.get_neighbour_matrix_by_hand(knn=3).T != 0
    ng, sched, nit, goc0, order = _schedule(sup, 1, 32)
    assert ng == 4
    for g in range(ng):
        blocks, first_done = 0, None
        for it in range(nit[g]):
            s = int(sched[g, it])
            blocks += bin((s >> 6) & 63).count("1")
            if first_done is None and (s >> 24) & 63:
                first_done = blocks
        assert first_done <= 0.6 * blocks, (g, first_done, blocks)


def _wgrad_units(support, FCK, FCN, NCK, NCN):
    lib = L.load()
    row = (C.c_uint32 * 17)(*[int(sum(1 << j for j in range(17) if support[i, j])) for i in range(17)])
    buf = (C.c_uint8 * (12 * 160))()
    lib.lcn_debug_tcw_units.restype = C.c_int
    n = lib.lcn_debug_tcw_units(row, FCK, FCN, NCK, NCN, buf, 160)
    return np.array(buf[:12 * n], dtype=np.uint8).reshape(n, 12)


@pytest.mark.parametrize("knn,FC", [(1, 1), (2, 1), (3, 1), (3, 2), (0, 1), (0, 2)])
def test_wgrad_units_cover_exactly_the_mask_blocks(knn, FC):
    """Weight-gradient units (pairs of input chunks x groups of <= 4 output chunks of the union of their neighbourhoods,
    lcn_gemm_tc.cu: tcw_build_units): every nonzero (input chunk, output chunk) block is stored by exactly one unit and
    row half, nothing outside the mask is stored, every input chunk sits in exactly one pair."""
    sup = (O.get_neighbour_matrix_by_hand(knn=knn).T != 0) if knn > 0 else np.ones((17, 17), bool)
    NC = 17 * FC
    units = _wgrad_units(sup, FC, FC, NC, NC)
    assert 0 < len(units) <= 160
    stored = np.zeros((NC, NC), int)
    partner = {}
    computed = 0
    for ic0, ic1, ln, _, *rest in units.tolist():
        ocs, keep = rest[:4], rest[4:]
        assert 1 <= ln <= 4 and ic0 < NC and (ic1 == 0xFF or ic1 < NC)
        partner.setdefault(ic0, ic1)
        assert partner[ic0] == ic1                      # an input chunk always appears with the same partner
        assert len(set(ocs[:ln])) == ln
        computed += ln * 2                              # M = 128 MMAs: both row halves run, single or not
        for q in range(ln):
            assert keep[q] & 3 and not keep[q] & ~3
            if keep[q] & 1:
                stored[ic0, ocs[q]] += 1
            if keep[q] & 2:
                assert ic1 != 0xFF
                stored[ic1, ocs[q]] += 1
    want = np.kron(sup, np.ones((FC, FC), int))
    assert np.array_equal(stored, want)
    firsts, seconds = set(partner), {v for v in partner.values() if v != 0xFF}
    assert not firsts & seconds and firsts | seconds == set(range(NC))
    if knn == 3 and FC == 1:
        assert len(units) == 31 and computed == 218     # the figures DESIGN.md quotes
    if FC == 2:
        assert computed == int(want.sum())              # the two halves of a joint share their neighbourhood: no waste


def test_wgrad_units_edge_layers():
    """Last layer (every input chunk -> the one padded output chunk) and first layer (one padded input chunk -> all)."""
    row1 = np.zeros((17, 17), bool)
    row1[:, 0] = True
    units = _wgrad_units(row1, 1, 1, 17, 1)
    assert len(units) == 9 and all(u[2] == 1 and u[4] == 0 for u in units.tolist())
    first = np.zeros((17, 17), bool)
    first[0, :] = True
    units = _wgrad_units(first, 1, 1, 1, 17)
    assert len(units) == 5 and all(u[0] == 0 and u[1] == 0xFF for u in units.tolist())
    assert sorted(oc for u in units.tolist() for oc in u[4:4 + u[2]]) == list(range(17))
