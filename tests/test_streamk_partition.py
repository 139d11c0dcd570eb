"""CPU: the work partition of the stream-K matrix-DFT kernel (csrc/mdft_tc.cu, tc_gemm_sk_kernel), restated in Python.

The kernel cuts tiles x chunks into gridDim.x contiguous ranges; a CTA walks its range as segments (tile, chunk range).
Invariants the device code relies on:
  * every chunk of every tile belongs to exactly one segment;
  * a segment that starts inside a tile is the FIRST segment of its CTA (so its partial is published before any wait) and
    a CTA has at most one such segment (one partial slot per CTA);
  * the head segment of a split tile is the LAST segment of its CTA and the pieces it gathers are exactly the first
    segments of CTAs g+1, g+2, ... up to the one that reaches the tile's end."""
import pytest


def segments(g, G, ntiles, upt):
    U = ntiles * upt
    u0, u1 = U * g // G, U * (g + 1) // G
    out, u = [], u0
    while u < u1:
        tile = u // upt
        ub = u - tile * upt
        ue = min(upt, ub + (u1 - u))
        out.append((tile, ub, ue))
        u += ue - ub
    return out


def gathered_by_owner(g, G, ntiles, upt, tile):
    U = ntiles * upt
    tile_u1 = (tile + 1) * upt
    got = []
    for gp in range(g + 1, G):
        got.append(gp)
        if U * (gp + 1) // G >= tile_u1:
            break
    return got


@pytest.mark.parametrize('ntiles,upt,G', [(128, 128, 148), (64, 32, 148), (16, 128, 148), (5, 7, 3), (33, 4, 132), (128, 128, 160), (8, 128, 148)])
def test_partition_invariants(ntiles, upt, G):
    cover = {}
    first_piece_owner = {}
    for g in range(G):
        segs = segments(g, G, ntiles, upt)
        mid_starts = [i for i, (t, ub, ue) in enumerate(segs) if ub > 0]
        assert mid_starts in ([], [0]), 'only the first segment of a CTA may start inside a tile'
        for i, (t, ub, ue) in enumerate(segs):
            assert 0 <= ub < ue <= upt
            for c in range(ub, ue):
                assert (t, c) not in cover
                cover[(t, c)] = g
            if ub == 0 and ue < upt:
                assert i == len(segs) - 1, 'the head of a split tile is the last segment of its CTA'
                first_piece_owner[t] = g
    assert len(cover) == ntiles * upt
    for t, g in first_piece_owner.items():
        pieces = gathered_by_owner(g, G, ntiles, upt, t)
        want = sorted({cover[(t, c)] for c in range(upt)} - {g})
        assert pieces == want
        for gp in pieces:   # each gathered CTA's partial is its first segment and lies in this tile
            s0 = segments(gp, G, ntiles, upt)[0]
            assert s0[0] == t and s0[1] > 0
