"""Host-side logic of the action-reaction (symmetric) R^3 path, no GPU needed: the rule builder must assign every
ordered (i-tile, j-tile) interaction of the whole job to exactly one evaluation -- one-sided inside an i-block,
or one symmetric evaluation that serves both directions (pair_r3_sym.cuh, engine.cu: build_sym_rules)."""
import numpy as np
import pytest

import steps_b200 as sb
from steps_b200 import _lib, api

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


def _schedule(n, nranks, rank, topology=0, real_bytes=8, budget=0):
    import ctypes as C

    lib = _lib.load()
    plan = (C.c_int * 8)()
    max_ctas, max_passes, max_words = 4_000_000, 4096, 8_000_000
    order = np.zeros(2 * max_ctas, dtype=np.int32)
    off = np.zeros(max_passes + 1, dtype=np.int32)
    mask = np.zeros(max_words, dtype=np.uint64)
    words = C.c_int()
    n_cta = lib.steps_b200_sym_schedule_host(n, nranks, rank, topology, real_bytes, budget, plan, order.ctypes.data_as(C.POINTER(C.c_int)), max_ctas,
                                             off.ctypes.data_as(C.POINTER(C.c_int)), max_passes, mask.ctypes.data_as(C.POINTER(C.c_ulonglong)),
                                             max_words, C.byref(words))
    assert n_cta >= 0, n_cta
    keys = ("ib_size", "n_ib", "sb", "n_sb", "n_chunks", "tpc", "n_tiles", "n_passes")
    p = dict(zip(keys, plan))
    return p, order[: 2 * n_cta].reshape(-1, 2), off[: p["n_passes"] + 1], mask[: p["n_ib"] * words.value].reshape(p["n_ib"], words.value)


def _block_chunk_activity(rules, p):
    """the kernel's own rule (pair_r3_sym.cuh sym_hull on the rule restricted to a chunk): a (block, chunk) is work -- and gets a
    partial sum written -- iff one of the block's tile ranges has a tile inside the chunk"""
    act = np.zeros((p["n_ib"], p["n_chunks"]), dtype=bool)
    for b, r in enumerate(rules):
        rngs = [(r[0], r[1])] + [(r[3 + k], r[8 + k]) for k in range(r[2])]
        for lo, hi in rngs:
            if lo < hi:
                act[b, lo // p["tpc"]: (hi - 1) // p["tpc"] + 1] = True
    return act


@pytest.mark.parametrize("n,nranks,topology,real_bytes", [(6001, 2, 0, 8), (6001, 3, 0, 8), (40000, 2, 0, 4), (200000, 8, 0, 8), (2_000_000, 1, 0, 8),
                                                         (2_000_000, 8, 0, 8), (262144, 4, 1, 8), (400000, 5, 3, 8), (16_777_216, 2, 0, 4)])
def test_launch_schedule_covers_exactly_what_the_final_reduction_reads(n, nranks, topology, real_bytes):
    """Host-only check of the action-reaction launch schedule (round 2): the chunk mask the final reduction reads, the kernel's own
    activity rule and the list of launched CTAs must agree for every rank -- a block's tile ranges can leave whole chunks out in between
    (ring assignment with gaps), which a contiguous chunk range got wrong and only a 2-GPU run with reused memory showed."""
    for rank in sorted({0, nranks // 2, nranks - 1}):
        p, order, off, mask = _schedule(n, nranks, rank, topology, real_bytes)
        _, _, rules = api.sym_rules(n, nranks, rank, p["ib_size"])
        rules = [[int(v) for v in r] for r in rules]
        assert len(rules) == p["n_ib"] and p["n_sb"] == -(-p["n_ib"] // p["sb"]) and p["n_chunks"] == -(-p["n_tiles"] // p["tpc"])
        act = _block_chunk_activity(rules, p)
        bits = np.array([[(int(mask[b, c >> 6]) >> (c & 63)) & 1 for c in range(p["n_chunks"])] for b in range(p["n_ib"])], dtype=bool)
        assert np.array_equal(bits, act), "chunk mask differs from the kernel's activity rule"
        # superblock x chunk combinations that carry work, pass by pass: launched exactly once, nothing else launched
        seen = set()
        n_pad = p["n_tiles"] * TJ
        rows = max(1, min((16 << 30) // (3 * n_pad * real_bytes), p["n_sb"]))  # superblock rows per pass, as the library sizes them
        assert p["n_passes"] == -(-p["n_sb"] // rows)
        for k in range(p["n_passes"]):
            ctas = order[off[k]: off[k + 1]]
            assert len(ctas) == 0 or int(ctas[:, 0].max()) < min(rows, p["n_sb"] - k * rows)
            for gs, jc in ctas:
                key = (k * rows + int(gs), int(jc))
                assert key not in seen, "a (superblock, chunk) combination is launched twice"
                seen.add(key)
        want = set()
        for S in range(p["n_sb"]):
            blk = act[S * p["sb"]: (S + 1) * p["sb"]]
            for jc in np.flatnonzero(blk.any(axis=0)):
                want.add((S, int(jc)))
        assert seen == want, "launched CTAs differ from the combinations that carry work"
        # every tile of every symmetric or diagonal range lies in an active (block, chunk): nothing is skipped
        for b, r in enumerate(rules):
            for lo, hi in [(r[0], r[1])] + [(r[3 + k], r[8 + k]) for k in range(r[2])]:
                for t in (lo, hi - 1):
                    if lo < hi:
                        assert act[b, t // p["tpc"]]


def test_launch_schedule_multi_pass_and_wave_count():
    """C5 on one GPU: the rows of all superblocks do not fit the buffer -> several passes, each with >= 100 waves' worth of working CTAs
    unless the chunk count is at its cap; C2 on one GPU: one pass"""
    p, order, off, _ = _schedule(16_777_216, 1, 0, 0, 4, 64 << 30)
    assert p["n_passes"] > 1 and p["sb"] == 8
    assert off[-1] == len(order)
    p2, order2, off2, _ = _schedule(2_000_000, 1, 0, 0, 8, 64 << 30)
    assert p2["n_passes"] == 1 and p2["sb"] == 8
    assert len(order2) >= 100 * 296  # >= 100 waves of 2 CTAs x 148 SMs
