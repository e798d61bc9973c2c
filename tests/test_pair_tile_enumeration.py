"""CPU check of the work decomposition of the pair-tile kernel (csrc/pair_forces.cu): tile pairs (I <= J) of 32 particles,
32 rotations (l, (l+t) mod 32) per off-diagonal tile pair, rotations 1..16 (16 on half the lanes) per diagonal tile, taken
TWO per loop step (pair_rotation2: even rotations of a run into one reaction accumulator, odd ones into the other), optionally
cut over `split` warps, items ordered tile pair by tile pair with the diagonal tiles last. Restated in plain Python with the
kernel's own index arithmetic and masks; the property: every unordered particle pair of every bead is evaluated exactly once,
and within one loop step the two reaction accumulators never see two lanes on the same slot.

Reference: the i<j double loop of Simulation::updatePhysicalForces (src/simulation.cpp:432-454)."""
import itertools

import pytest

TILE = 32


def tile_list(T):
    """api.cu: off-diagonal tile pairs first (row-major), the diagonal ones last."""
    return [(i, j) for i in range(T) for j in range(i + 1, T)] + [(i, i) for i in range(T)]


def enumerate_pairs(N, nbeads, split):
    """Every (bead, i, j) the kernel evaluates, with the masks of pair_rotation2 (MASKED variant; the unmasked variant is the
    same loop on full off-diagonal tiles, where every mask is true)."""
    T = (N + TILE - 1) // TILE
    tiles = tile_list(T)
    TP = len(tiles)
    out = []
    nwarps = nbeads * TP * split
    for gw in range(nwarps):
        item, part = divmod(gw, split)
        bl = item % nbeads                      # items run tile pair by tile pair: all beads of one pair are neighbours
        I, J = tiles[item // nbeads]
        diag = I == J
        nrot = (16 if diag else 32) // split
        tb = (1 if diag else 0) + part * nrot
        assert nrot % 2 == 0
        for t in range(tb, tb + nrot, 2):
            slots_a, slots_b = set(), set()
            for lane in range(TILE):
                pi = I * TILE + lane
                vi = pi < N
                sa, sb = (lane + t) % TILE, (lane + t + 1) % TILE
                slots_a.add(sa); slots_b.add(sb)
                acta = vi and (J * TILE + sa < N) and not (diag and t == 16 and lane >= 16)
                actb = vi and (J * TILE + sb < N) and not (diag and t + 1 == 16 and lane >= 16)
                if acta:
                    out.append((bl, pi, J * TILE + sa))
                if actb:
                    out.append((bl, pi, J * TILE + sb))
            assert len(slots_a) == TILE and len(slots_b) == TILE     # conflict-free read-modify-writes in each accumulator
    return out


@pytest.mark.parametrize("N,nbeads,split", [(64, 2, 1), (70, 3, 1), (33, 1, 1), (96, 2, 2), (100, 1, 4), (31, 2, 1), (128, 1, 2)])
def test_every_unordered_pair_exactly_once(N, nbeads, split):
    got = enumerate_pairs(N, nbeads, split)
    norm = sorted((b, min(i, j), max(i, j)) for b, i, j in got)
    assert all(i != j for _, i, j in norm)
    want = sorted((b, i, j) for b in range(nbeads) for i, j in itertools.combinations(range(N), 2))
    assert norm == want


def test_diagonal_tiles_come_last_and_are_half_size():
    tiles = tile_list(6)
    assert len(tiles) == 6 * 7 // 2
    assert all(i < j for i, j in tiles[:15]) and all(i == j for i, j in tiles[15:])
