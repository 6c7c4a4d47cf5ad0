"""Host-side logic of the action-reaction (symmetric) R^3 path, no GPU needed: the rule builder must assign every
ordered (i-tile, j-tile) interaction of the whole job to exactly one evaluation -- one-sided inside an i-block,
or one symmetric evaluation that serves both directions (pair_r3_sym.cuh, engine.cu: build_sym_rules)."""
import numpy as np
import pytest

import steps_b200 as sb
from steps_b200 import api

TJ = 128


def _coverage(n, nranks, ib):
    n_tiles = (n + TJ - 1) // TJ
    tpb = ib // TJ
    cover = np.zeros((n_tiles, n_tiles), dtype=np.int32)  # cover[u, v]: times the forces of v's particles ON u's particles are formed
    rows = np.zeros(n_tiles, dtype=np.int32)              # ownership of tiles as i-rows
    prev_hi = 0
    for r in range(nranks):
        res = api.sym_rules(n, nranks, r, ib)
        assert res is not None
        lo, hi, rules = res
        assert lo == prev_hi and lo % ib == 0, "partition must be contiguous and i-block aligned"
        prev_hi = hi
        for b, ru in enumerate(rules):
            d_lo, d_hi, n_sym = int(ru[0]), int(ru[1]), int(ru[2])
            assert d_lo == (lo // TJ) + b * tpb and d_lo < d_hi <= n_tiles
            rows[d_lo:d_hi] += 1
            own = slice(d_lo, d_hi)
            cover[own, d_lo:d_hi] += 1  # one-sided: own tiles act on own tiles (both directions appear as separate (u,v))
            last = -1
            for k in range(n_sym):
                s_lo, s_hi = int(ru[3 + k]), int(ru[8 + k])
                assert 0 <= s_lo < s_hi <= n_tiles and s_lo > last, "ranges sorted, disjoint, non-empty"
                last = s_hi - 1
                assert s_hi <= d_lo or s_lo >= d_hi, "symmetric range must not overlap the block's own tiles"
                cover[own, s_lo:s_hi] += 1      # F_i += ...
                cover[s_lo:s_hi, own] += 1      # F_j -= ...
    assert prev_hi == n
    assert (rows == 1).all()
    return cover


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("n,ib", [(20000, 768), (33333, 1024), (100000, 896), (16 * 768, 768), (8000, 384), (2744, 256), (5000, 128), (40000, 2048), (30000, 1536)])
def test_every_interaction_exactly_once(n, nranks, ib):
    nb_total = (n + ib - 1) // ib
    if nb_total < 2 * nranks:
        assert api.sym_rules(n, nranks, 0, ib) is None  # too few blocks: the engine stays on the one-sided kernel
        return
    cover = _coverage(n, nranks, ib)
    assert cover.min() == 1 and cover.max() == 1


@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_work_is_balanced(nranks):
    """symmetric evaluations are the cost unit (20 FP64 instructions each); one-sided diagonal blocks are negligible"""
    n, ib = 2_000_000, 768
    tpb = ib // TJ
    work = []
    for r in range(nranks):
        lo, hi, rules = api.sym_rules(n, nranks, r, ib)
        w = 0
        for ru in rules:
            for k in range(int(ru[2])):
                w += (int(ru[8 + k]) - int(ru[3 + k])) * tpb
        work.append(w)
    work = np.array(work, dtype=float)
    assert work.max() / work.mean() < 1.01
    # and the total is half of the one-sided work
    n_tiles = (n + TJ - 1) // TJ
    nb = (n + ib - 1) // ib
    assert abs(work.sum() / (0.5 * n_tiles * nb * tpb) - 1) < 0.01


def test_rejects_bad_geometry():
    assert api.sym_rules(700, 1, 0, 768) is None        # fewer than two i-blocks
    assert api.sym_rules(100000, 4, 0, 700) is None     # i-block not a multiple of the j-tile


def test_chunk_target_of_the_action_reaction_plan():
    """steps_b200_sym_chunk_target (host-only): 56 chunks for everything measured in round 1, more and shorter chunks only when a
    pass of a very large problem would hold fewer than 24 waves of CTAs (tools/pass_model.py)"""
    from steps_b200 import _lib

    f = _lib.load().steps_b200_sym_chunk_target
    sms = 148
    # C2: N = 2M FP64, i-block 768, 2 CTAs/SM, 341 rows of 48 MB in 16 GB
    assert f(15625, 2605, 341, sms * 2) == 56
    # C2 on 8 GPUs: 326 own i-blocks per rank
    assert f(15625, 326, 341, sms * 2) == 56
    # FP32 at N = 2M (i-block 1024, 4 CTAs/SM): every row fits
    assert f(15625, 1954, 1954, sms * 4) == 56
    # small problems never change
    assert f(3125, 521, 521, sms * 2) == 56 and f(71, 12, 1, sms * 2) == 56
    # C5: N = 16.7M FP32, 79 rows of 201 MB per pass -> capped at 160 chunks
    assert f(131072, 16384, 79, sms * 4) == 160
    # C4 with the S^1xR^2 action-reaction kernel: i-block 384, 160 rows of 100 MB
    assert f(32768, 10923, 160, sms * 4) == 89
    # monotone in the rows a pass holds
    vals = [f(131072, 16384, r, sms * 4) for r in (40, 79, 160, 320, 640)]
    assert vals == sorted(vals, reverse=True) and vals[-1] == 56
