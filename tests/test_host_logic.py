"""Host-side logic that needs no GPU: random streams, sharding, file readers, logging schedule."""
import os

import numpy as np
import pytest

from ppde_b200 import dist as D
from ppde_b200 import philox, weights
from ppde_b200.synthetic import synthetic_problem


def test_philox_random123_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_streams_are_indexed_by_global_chain():
    u_all = philox.proposal_uniforms(9, 3, 1, np.arange(8), 50)
    u_sh = philox.proposal_uniforms(9, 3, 1, np.arange(4, 8), 50)
    assert np.array_equal(u_all[4:], u_sh)
    assert u_all.min() > 0 and u_all.max() < 1
    U = philox.path_lengths(9, 3, np.arange(1000), 2)
    assert set(np.unique(U)) == {1, 2, 3}
    U10 = philox.path_lengths(9, 3, np.arange(4000), 10)
    assert U10.min() == 1 and U10.max() == 19


def test_shard_ranges_cover_population():
    for n in (1, 7, 128, 65536, 65537):
        for ws in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == D.shard_sizes(n, ws)
            if n >= ws:
                assert D.owner_of(n - 1, n, ws) == (ws - 1, sizes[-1] - 1)


def test_quantile_rule_of_the_report_kernel_is_numpys():
    """quantile_kernel (csrc/population.cu) selects the order statistics lo = floor(q (n-1)), lo + 1 and interpolates in
    double with numpy's _lerp rule (a + (b-a) t, or b - (b-a)(1-t) for t >= 0.5): restated here on a sorted array and
    compared with np.quantile (ppde.py:158-160 uses q = 0.5, 0.9) for many n, including ties and n = 1."""
    def kernel_rule(x, q):
        xs = np.sort(x.astype(np.float32))
        n = len(xs)
        v = q * (n - 1)
        lo = int(np.floor(v)); hi = min(lo + 1, n - 1); t = v - lo
        a, b = float(xs[lo]), float(xs[hi])
        return b - (b - a) * (1 - t) if t >= 0.5 else a + (b - a) * t
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 11, 128, 1001, 65536):
        x = rng.standard_normal(n).astype(np.float32)
        x[rng.integers(0, n, size=n // 3)] = 0.25
        for q in (0.5, 0.9, 0.0, 1.0, 0.2):
            # numpy evaluates the interpolation in the input dtype (float32), the kernel in double: agreement to float32 rounding
            assert kernel_rule(x, q) == pytest.approx(float(np.quantile(x, q)), rel=1e-6, abs=1e-7), (n, q)


def test_fasta_reader_and_offset_rule(tmp_path):
    f = tmp_path / "wt.fasta"
    f.write_text(">PABP_YEAST/115-210 some description\nQRDPS\nLRKKG\n>second\nAC\n")
    seqs, ids = weights.read_fasta(str(f))
    assert seqs == ["QRDPSLRKKG", "AC"] and ids == ["PABP_YEAST/115-210", "second"]
    assert weights.fasta_offset(ids[0]) == 115 and weights.fasta_offset("sarkisyan_wt") == 1   # nets.py:257-261
    aa = weights.seq_to_aa("ACDEFGHIKLMNPQRSTVWY")
    assert list(aa) == list(range(20)) and weights.aa_to_seq(aa) == "ACDEFGHIKLMNPQRSTVWY"
    with pytest.raises(ValueError):
        weights.seq_to_aa("AXB")


def test_potts_pickle_reader(tmp_path):
    import pickle
    pr = synthetic_problem(10, window=(2, 7))
    with open(tmp_path / "potts.pkl", "wb") as fh:
        pickle.dump({"J_ij": pr["J"], "h_i": pr["h"], "index_list": np.arange(2, 8) + 115, "reg_coef": 2.5}, fh)
    got = weights.load_potts(str(tmp_path), "PABP_YEAST/115-210")
    assert got["win_lo"] == 2 and got["win_hi"] == 7 and got["reg_coef"] == 2.5
    assert got["J"].shape == (6, 6, 20, 20)


def test_sampler_log_schedule():
    """Logging fires after iteration i when i > 0 and (i+1) % log_every == 0 (ppde.py:155)."""
    def fired(num_steps, log_every):
        out, t = [], 0
        while t < num_steps:
            i = max(t, 1)
            i += (-(i + 1)) % log_every
            stop = min(i + 1, num_steps)
            t = stop
            if t == i + 1:
                out.append(i)
        return out
    for T, le in ((100, 50), (10, 3), (7, 1), (5, 10)):
        want = [i for i in range(T) if i > 0 and (i + 1) % le == 0]
        assert fired(T, le) == want


def test_sampler_constructor_flag_semantics():
    import argparse
    from ppde_b200.sampler import PPDE_PAS
    s = PPDE_PAS(argparse.Namespace(ppde_pas_length=3, nmut_threshold=0, paper_results=True))
    assert s.nmut_threshold == np.iinfo(np.int32).max and s.ppde_temp == 2 and s.paper_results   # ppde.py:9-17
    assert s.approximate_energy_change(4.0) == 2.0


def test_block_table_slot_rule_never_overwrites_a_live_slot():
    """Host restatement of the pool-row block tables of the incremental CNN forward (cnn_inc_merge_kernel /
    cnn_forward_inc_kernel in ppde_b200/csrc/cnn_tc.cu).  A chain owns two private rows (b, n + b) and may point at
    read-only fixed rows; a proposal row Y built from the current row X takes, for a dirty block q, Y's own slot unless
    X's table points at it - then X's own slot.  Whatever the sequence of accepts, rejects and resets to a fixed row,
    (a) the slot written is referenced by neither the current row's table nor a fixed row, and (b) reading a live row
    through its table returns exactly the blocks of the state it represents.  (Inside the step loop the forward is always
    incremental; full evaluations only fill rows at t = 0, where every row points at itself.)"""
    rng = np.random.default_rng(7)
    NB, FIX = 7, 2                      # rows: 0 = A, 1 = B (private), 2.. = fixed
    A, B = 0, 1
    for trial in range(200):
        content = {}                     # (row, q) -> version stored in that slot
        tab = {}                         # row -> source row of every block
        truth = {}                       # row -> versions of the state the row represents
        ver = 0
        for f in range(FIX):             # fixed rows: fully evaluated, point at themselves
            r = 2 + f
            tab[r] = [r] * NB
            truth[r] = []
            for q in range(NB):
                ver += 1; content[(r, q)] = ver; truth[r].append(ver)
        cur = 2                          # the chain starts on a fixed row (wild type)
        for step in range(80):
            Y = A if cur != A else B     # engine rule: the private row that is not the current one (ppde::y_row)
            X = cur
            dirty = [bool(rng.random() < 0.3) for _ in range(NB)]
            new_tab, new_truth = [], []
            for q in range(NB):
                if dirty[q]:
                    target = X if tab[X][q] == Y else Y
                    assert target in (A, B), "fixed rows are read-only"
                    assert tab[X][q] != target, "the slot written is still referenced by the current row"
                    ver += 1
                    content[(target, q)] = ver
                    new_tab.append(target); new_truth.append(ver)
                else:
                    new_tab.append(tab[X][q]); new_truth.append(truth[X][q])
            tab[Y], truth[Y] = new_tab, new_truth
            for r in (X, Y):             # both live rows read back correctly through their tables
                assert [content[(tab[r][q], q)] for q in range(NB)] == truth[r], (trial, step, r)
            u = rng.random()
            if u < 0.7:
                cur = Y                  # accept
            elif u < 0.85:
                cur = 2 + int(rng.integers(FIX))   # hard reset / paper-mode reject: back to a fixed row
            # else: reject, keep X


def test_exact_backward_schedule():
    """Delta backward with a periodic exact refresh: which iterations run the exact backward."""
    from ppde_b200.engine import ChainEngine

    class _M:                            # the two attributes full_backward_at reads
        bwd_refresh = 32

    e = ChainEngine.__new__(ChainEngine)
    e.m = _M()
    e.delta = True
    full = [t for t in range(100) if e.full_backward_at(t)]
    assert full == [31, 63, 95]
    e.delta = False
    assert all(e.full_backward_at(t) for t in range(5))


def test_delta_record_work_order_is_consistent():
    """The three places that permute the columns of a delta-backward tile must agree (csrc/cnn_tc.cu): the record builder ranks
    the entries by (tile, producer warp w = r % NW, slot sl = r // NW) with r the column inside the tile, a producer warp stores
    the column of its slot sl into operand row w * RPW + sl, and the epilogue maps accumulator column j back to tile column
    (j % RPW) * NW + j // RPW.  Host-side restatement with the kernel's constants (BD_NT = 48, BD_NW = 6, BD_RPW = 8); also the
    per-warp offsets as 'first rank of the next non-empty group' (suffix minimum)."""
    NT, NW = 48, 6
    RPW = NT // NW
    rng = np.random.default_rng(0)
    for npos in (0, 1, 5, 47, 48, 49, 130, 252):
        ntile = (npos + NT - 1) // NT
        cols = np.arange(npos)
        t, r = cols // NT, cols % NT
        w, sl = r % NW, r // NW
        g = t * NT + w * RPW + sl                               # work-order key of a column
        assert len(set(g.tolist())) == npos                      # a permutation: no two columns share a place
        row = w * RPW + sl                                      # operand row inside the tile
        back = (row % RPW) * NW + row // RPW                    # the epilogue's inverse
        assert np.array_equal(back, r)
        # entries: a few per column; sorted by (g, side, channel) they must be contiguous per (tile, warp) group, slots ascending
        ne = rng.integers(0, 4, size=npos)
        ent = [(int(g[c]), int(s), int(ch)) for c in range(npos) for s, ch in zip(rng.integers(0, 2, ne[c]), rng.choice(476, ne[c], replace=False))]
        ent.sort()
        grp = np.array([(NW * (e[0] // NT) + (e[0] % NT) // RPW) for e in ent], dtype=int)
        nw = NW * ntile + 1
        first = np.full(nw, 2 ** 31 - 1)
        for rank, q in enumerate(grp):
            first[q] = min(first[q], rank)
        first[-1] = min(first[-1], len(ent))
        woff = np.minimum.accumulate(np.where(first == 2 ** 31 - 1, len(ent), first)[::-1])[::-1]
        assert woff[0] == 0 or len(ent) == 0 or grp.min() > 0
        for q in range(nw - 1):
            seg = [e for e, gq in zip(ent, grp) if gq == q]
            assert woff[q + 1] - woff[q] == len(seg)
            slots = [(e[0] % NT) % RPW for e in seg]
            assert slots == sorted(slots)
