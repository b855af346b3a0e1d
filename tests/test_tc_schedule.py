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
