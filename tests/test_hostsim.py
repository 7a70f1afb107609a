"""CPU checks of the product's host code and of the kernel's search state machine compiled for the host
(tests/hostsim): index builder vs the oracle's naive suffix sort, counts vs the oracle / brute force."""
import os

import numpy as np
import pytest

import gmtest as T


@pytest.mark.parametrize("seed,nchr,length", [(1, 1, 400), (2, 3, 1500), (3, 5, 77), (4, 2, 20000)])
def test_index_builder_matches_oracle_suffix_sort(seed, nchr, length):
    seqs = T.repeat_rich(seed, nchr, length)
    orc, hs = T.Oracle(seqs), T.HostSim(seqs, with_sa=True)
    for rev in (False, True):
        a, b = orc.bwt(rev), hs.bwt(rev)
        assert np.array_equal(np.minimum(a, 5), b), "BWT differs (rev=%s)" % rev
    assert np.array_equal(orc.sa().astype(np.uint32), hs.sa())


def test_index_builder_degenerate_texts():
    for seqs in ([np.zeros(500, np.uint8)], [np.tile(np.array([0, 1], np.uint8), 300)] * 3,
                 [np.array([2], np.uint8), np.array([2], np.uint8)], [np.full(193, 3, np.uint8), np.full(191, 3, np.uint8)]):
        orc, hs = T.Oracle(seqs), T.HostSim(seqs, with_sa=True)
        for rev in (False, True):
            assert np.array_equal(orc.bwt(rev), hs.bwt(rev))
        assert np.array_equal(orc.sa().astype(np.uint32), hs.sa())


@pytest.mark.parametrize("K,E", [(30, 0), (30, 1), (30, 2), (21, 3), (16, 4), (50, 2), (12, 2), (9, 1), (8, 0),
                                 (33, 1), (64, 2), (65, 1), (2, 0), (3, 1), (4, 2), (5, 3), (6, 4), (130, 3)])
def test_state_machine_matches_oracle(K, E):
    seqs = T.repeat_rich(7, 3, 3000)
    orc, hs = T.Oracle(seqs), T.HostSim(seqs)
    for rc in (True, False):
        want = orc.map(K, E, revcompl=rc)
        got = hs.map(K, E, revcompl=rc)
        assert np.array_equal(got, want), (K, E, rc, np.nonzero(got != want)[0][:10])


@pytest.mark.parametrize("K,E", [(30, 0), (30, 2), (21, 3), (16, 4), (50, 2), (8, 0), (3, 1), (65, 1)])
def test_jump_table_depths_do_not_change_results(K, E):
    seqs = T.repeat_rich(17, 2, 2500)
    hs = T.HostSim(seqs)
    base = hs.map(K, E, jump_depth=0)
    assert np.array_equal(base, T.Oracle(seqs).map(K, E))
    for d in (1, 2, 5, 7, 8):
        assert np.array_equal(hs.map(K, E, jump_depth=d), base), (K, E, d)
        assert np.array_equal(hs.map(K, E, jump_depth=d, revcompl=False), hs.map(K, E, jump_depth=0, revcompl=False))


def test_state_machine_saturation_and_palindromes():
    seqs = [np.tile(np.array([0, 1, 2, 3], dtype=np.uint8), 2000)]
    orc, hs = T.Oracle(seqs), T.HostSim(seqs)
    for bits in (8, 16):
        want = orc.map(8, 0, value_bits=bits)
        assert np.array_equal(hs.map(8, 0, value_bits=bits), want)
        assert want.max() == (255 if bits == 8 else 3998)
    assert np.array_equal(hs.map(10, 2, value_bits=8), orc.map(10, 2, value_bits=8))


def test_state_machine_short_sequences_and_selection():
    seqs = [np.array([0, 1, 2], np.uint8), T.repeat_rich(3, 1, 300)[0], np.array([3, 3], np.uint8),
            T.repeat_rich(4, 1, 200)[0], np.array([2], np.uint8)]
    orc, hs = T.Oracle(seqs), T.HostSim(seqs)
    assert np.array_equal(hs.map(12, 1), orc.map(12, 1))
    assert np.array_equal(hs.map(12, 1), T.brute(seqs, 12, 1))
    iv = [(0, 2), (10, 40), (35, 60), (290, 320), (500, 506)]
    assert np.array_equal(hs.map(12, 1, intervals=iv), orc.map(12, 1, intervals=iv))
    # sharding: disjoint position ranges add up to the whole
    whole = hs.map(12, 1)
    parts = sum(hs.map(12, 1, pos_begin=b, pos_end=e).astype(np.int64) for b, e in [(0, 100), (100, 333), (333, 506)])
    assert np.array_equal(parts, whole)


def test_multi_file_counts_whole_index():
    base = T.repeat_rich(13, 2, 1200)
    rng = np.random.default_rng(5)
    seqs, stf = [], []
    for g in range(3):
        for s in base:
            s = s.copy()
            m = rng.random(len(s)) < 0.02 * g
            s[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
            seqs.append(s); stf.append(g)
    stf = np.array(stf, dtype=np.uint32)
    orc, hs = T.Oracle(seqs, seq_to_file=stf), T.HostSim(seqs)
    for f in range(3):
        assert np.array_equal(hs.map(25, 2, seq_to_file=stf, file_no=f), orc.map(25, 2, file_no=f))


def test_fetch_counter_is_deterministic_and_plausible():
    seqs = T.repeat_rich(21, 2, 4000)
    hs = T.HostSim(seqs)
    out, f1 = hs.map(30, 0, return_fetches=True)
    _, f2 = hs.map(30, 0, return_fetches=True)
    assert f1 == f2
    n_kmers = sum(len(s) - 29 for s in seqs)
    _, f0 = hs.map(30, 0, return_fetches=True, jump_depth=0)
    assert n_kmers * 10 < f0 < n_kmers * 130  # <= 30 steps per strand, 1-2 blocks per step
    assert f1 < f0  # the jump table skips the top levels


@pytest.mark.parametrize("case", ["3c", "3d", "3e", "3f"])
def test_exclude_pseudo_matches_reference_golden(case):
    cfg = T.CASES[case]
    files, sel, folder = T.load_case(case)
    seqs, stf, _ = T.case_layout(files)
    hs = T.HostSim(seqs, with_sa=True)
    for fi, (base, recs) in enumerate(files):
        iv = T.file_intervals(sel, recs)
        if iv is None:
            continue
        import os
        gold = np.fromfile(os.path.join(folder, "raw_freq16", base + ".genmap.freq16"), dtype=np.uint16)
        got = hs.map(cfg["K"], cfg["E"], revcompl=cfg["rc"], seq_to_file=stf, file_no=fi, intervals=iv, exclude_pseudo=True)
        assert np.array_equal(got, gold), (case, base)


@pytest.mark.parametrize("K,E,rc", [(25, 2, True), (25, 1, False), (12, 0, True), (40, 3, True)])
def test_exclude_pseudo_matches_oracle(K, E, rc):
    base = T.repeat_rich(13, 2, 1500)
    rng = np.random.default_rng(5)
    seqs, stf = [], []
    for g in range(4):
        for s in base:
            s = s.copy()
            m = rng.random(len(s)) < 0.03 * g
            s[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
            seqs.append(s); stf.append(g)
    stf = np.array(stf, dtype=np.uint32)
    orc, hs = T.Oracle(seqs, seq_to_file=stf), T.HostSim(seqs, with_sa=True)
    for f in (0, 3):
        want = orc.map(K, E, revcompl=rc, exclude_pseudo=True, file_no=f)
        for depth in (0, -1):
            got = hs.map(K, E, revcompl=rc, seq_to_file=stf, file_no=f, exclude_pseudo=True, jump_depth=depth)
            assert np.array_equal(got, want), (K, E, rc, f, depth)


@pytest.mark.parametrize("K,E", [(30, 0), (30, 1), (30, 2), (21, 3), (16, 4), (50, 2), (9, 1), (3, 1), (4, 2), (65, 1), (33, 2)])
def test_block_size_does_not_change_results(K, E):
    """The reference's -xo invariance (tests/tests.sh:47-60): any number of k-mers per block gives the same counts."""
    seqs = T.repeat_rich(7, 3, 2000) + [np.array([0, 1, 2], np.uint8), T.repeat_rich(9, 1, K + 3)[0]]
    orc, hs = T.Oracle(seqs), T.HostSim(seqs)
    want = orc.map(K, E)
    want_nc = orc.map(K, E, revcompl=False)
    for B in (1, 2, 3, 5, 8, 16, 0):
        for depth in (0, -1):
            assert np.array_equal(hs.map(K, E, block_kmers=B, jump_depth=depth), want), (K, E, B, depth)
        assert np.array_equal(hs.map(K, E, block_kmers=B, revcompl=False), want_nc), (K, E, B)


def test_blocked_search_saturation_selection_and_exclude_pseudo():
    pal = [np.tile(np.array([0, 1, 2, 3], dtype=np.uint8), 1500)]
    po, ph = T.Oracle(pal), T.HostSim(pal)
    for bits in (8, 16):
        for B in (1, 4, 7):
            assert np.array_equal(ph.map(10, 2, value_bits=bits, block_kmers=B), po.map(10, 2, value_bits=bits))
    seqs = T.repeat_rich(3, 2, 400)
    orc, hs = T.Oracle(seqs), T.HostSim(seqs)
    iv = [(0, 2), (10, 41), (35, 61), (390, 420), (700, 800)]
    for B in (1, 3, 6):
        assert np.array_equal(hs.map(12, 1, intervals=iv, block_kmers=B), orc.map(12, 1, intervals=iv))
        parts = sum(hs.map(12, 1, pos_begin=b, pos_end=e, block_kmers=B).astype(np.int64) for b, e in [(0, 101), (101, 333), (333, 800)])
        assert np.array_equal(parts, orc.map(12, 1))
    base = T.repeat_rich(13, 2, 900)
    rng = np.random.default_rng(5)
    ms, stf = [], []
    for g in range(4):
        for s in base:
            s = s.copy()
            m = rng.random(len(s)) < 0.03 * g
            s[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
            ms.append(s); stf.append(g)
    stf = np.array(stf, dtype=np.uint32)
    mo, mh = T.Oracle(ms, seq_to_file=stf), T.HostSim(ms, with_sa=True)
    for f in (0, 2):
        want = mo.map(25, 2, exclude_pseudo=True, file_no=f)
        for B in (1, 4, 6):
            assert np.array_equal(mh.map(25, 2, seq_to_file=stf, file_no=f, exclude_pseudo=True, block_kmers=B), want), (f, B)


# ---- Dna5: genomes containing N (src/algo.hpp:111-112,148-149: a pattern N never matches, a text N costs an error) ----
def test_dna5_index_builder_matches_oracle():
    for seed, nchr, length in [(11, 3, 800), (12, 1, 5000)]:
        seqs = T.repeat_rich(seed, nchr, length, with_n=True)
        seqs[0][10:40] = 4  # a run of N
        orc, hs = T.Oracle(seqs), T.HostSim(seqs, with_sa=True)
        assert orc.sigma == 5
        for rev in (False, True):
            assert np.array_equal(orc.bwt(rev), hs.bwt(rev))
        assert np.array_equal(orc.sa().astype(np.uint32), hs.sa())


@pytest.mark.parametrize("case", ["1c", "1d", "1e", "1f", "1g"])
def test_dna5_matches_reference_golden(case):
    import os
    cfg = T.CASES[case]
    files, sel, folder = T.load_case(case)
    seqs, stf, _ = T.case_layout(files)
    hs = T.HostSim(seqs)
    for bits, ext in ((16, "freq16"), (8, "freq8")):
        gold = np.fromfile(os.path.join(folder, "raw_" + ext, "genome.genmap." + ext), dtype=np.uint16 if bits == 16 else np.uint8)
        iv = T.file_intervals(sel, files[0][1])
        for B in (1, 0):
            got = hs.map(cfg["K"], cfg["E"], revcompl=cfg["rc"], value_bits=bits, intervals=iv, block_kmers=B)
            assert np.array_equal(got, gold), (case, bits, B)


@pytest.mark.parametrize("K,E", [(20, 0), (20, 1), (20, 2), (14, 3), (9, 4), (30, 2), (40, 1)])
def test_dna5_matches_oracle(K, E):
    seqs = T.repeat_rich(11, 3, 2500, with_n=True)
    seqs[1][100:160] = 4
    seqs[2][-5:] = 4
    orc, hs = T.Oracle(seqs), T.HostSim(seqs)
    for rc in (True, False):
        want = orc.map(K, E, revcompl=rc)
        for B in (1, 3, 0):
            for depth in (0, -1):
                got = hs.map(K, E, revcompl=rc, block_kmers=B, jump_depth=depth)
                assert np.array_equal(got, want), (K, E, rc, B, depth, np.nonzero(got != want)[0][:10])


@pytest.mark.parametrize("K,E,B,depth", [(20, 1, 0, -1), (20, 1, 1, -1), (20, 2, 4, 5), (14, 3, 0, -1), (9, 4, 2, -1), (33, 2, 3, -1), (40, 1, 0, -1)])
def test_dna5_searches_that_skip_the_text_n_plus_the_n_pass(K, E, B, depth, monkeypatch):
    """Dna5 index WITH the suffix array, E >= 1: the searches never descend into an N child (so they enter through
    substituted keys like on a Dna4 index; two-phase driver and general state machine), and the alignments to the text
    windows with 1..E N are added by the N pass — here its brute-force mirror.  Gap edges, a short run, isolated N, N at
    a sequence end; both strands, selection intervals, 8-bit saturation; against the oracle."""
    seqs = T.repeat_rich(51, 3, 2500, with_n=True)
    seqs[1][100:160] = 4
    seqs[1][300:302] = 4
    seqs[2][-5:] = 4
    seqs[0][[7, 500, 501, 1200, 1230]] = 4
    orc, hs = T.Oracle(seqs), T.HostSim(seqs, with_sa=True)
    walked = T.HostSim(seqs)  # no suffix array: the N children are walked, no substituted keys
    for env in ({}, {"GMB_BLOCK_KERNEL": "0"}):
        monkeypatch.delenv("GMB_BLOCK_KERNEL", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for rc in (True, False):
            want = orc.map(K, E, revcompl=rc)
            got = hs.map(K, E, revcompl=rc, block_kmers=B, jump_depth=depth)
            assert np.array_equal(got, want), (K, E, B, depth, rc, env, np.nonzero(got != want)[0][:10])
    if (K, E, B, depth) == (20, 2, 4, 5):
        keys = hs.last_lut_reads
        walked.map(K, E, revcompl=False, block_kmers=B, jump_depth=depth)
        assert keys > walked.last_lut_reads  # the substituted keys were in use
    iv = np.array([[90, 400], [2400, 2600], [5100, 7400]], dtype=np.uint64)
    want = orc.map(K, E, value_bits=8, intervals=iv)
    assert np.array_equal(hs.map(K, E, value_bits=8, intervals=iv, block_kmers=B, jump_depth=depth), want)
    monkeypatch.setenv("GMB_DNA5_NFREE", "0")
    assert np.array_equal(hs.map(K, E, block_kmers=B, jump_depth=depth), orc.map(K, E))


def test_pattern_n_mask_queries_for_every_offset_and_length():
    """Pattern<KW, 5>::has_n_in (any length: the common infix of a block), has_n(a, d <= 16) (table keys) and has_n()
    against a per-character loop — a shift by 32 or more is where x86 and the GPU part ways (profiles/r02 s28)."""
    assert T.hostsim_lib().hs_has_n_selftest(12345) == 0


def test_dna5_exclude_pseudo_and_reference_fixtures():
    import test_ref_fixtures as RF
    for line in [c for c in RF.CASES if c.startswith("dna5")]:
        name, K, E, flags, bits, seqs, stf, outs = RF.load_fixture(line)
        hs = T.HostSim(seqs)
        assert np.array_equal(hs.map(K, E, value_bits=bits), outs[0]), name
    base = T.repeat_rich(13, 2, 900, with_n=True)
    rng = np.random.default_rng(5)
    ms, stf = [], []
    for g in range(3):
        for s in base:
            s = s.copy()
            m = (rng.random(len(s)) < 0.03 * g) & (s < 4)
            s[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
            ms.append(s); stf.append(g)
    stf = np.array(stf, dtype=np.uint32)
    mo, mh = T.Oracle(ms, seq_to_file=stf), T.HostSim(ms, with_sa=True)
    for f in (0, 2):
        want = mo.map(22, 2, exclude_pseudo=True, file_no=f)
        for B in (1, 4):
            assert np.array_equal(mh.map(22, 2, seq_to_file=stf, file_no=f, exclude_pseudo=True, block_kmers=B), want), (f, B)


# ---- locate instantiation (csv lists, src/algo.hpp:311-346) ---------------------------------------------------
@pytest.mark.parametrize("with_n", [False, True])
@pytest.mark.parametrize("K,E,rc", [(12, 0, True), (12, 1, True), (14, 2, True), (10, 2, False), (33, 3, True)])
def test_locate_lists_match_definition(K, E, rc, with_n):
    seqs = T.repeat_rich(7, 3, 700, with_n=with_n)
    hs = T.HostSim(seqs, with_sa=True)
    _, limits = T.concat(seqs)
    got = hs.locate(K, E, revcompl=rc)
    want = T.brute_locations(seqs, K, E, T.valid_starts(limits, K), revcompl=rc)
    assert sum(len(a) + len(b) for a, b in got.values()) > int(limits[-1]) // 2
    for j in range(int(limits[-1])):
        assert got[j] == want.get(j, ([], [])), j
    for depth in (0, 3):  # jump tables do not change the lists
        assert hs.locate(K, E, revcompl=rc, jump_depth=depth) == got


@pytest.mark.parametrize("case", sorted(T.CASES))
def test_locate_lists_reproduce_reference_golden_csv(case):
    """All 18 csv goldens (tests/CMakeLists.txt:52-53, `-d`), selection and multi-FASTA cases included."""
    cfg = T.CASES[case]
    files, sel, folder = T.load_case(case)
    seqs, stf, _ = T.case_layout(files)
    hs = T.HostSim(seqs, with_sa=True)
    names = [f + ".fa" for f, _ in files]
    last = (np.cumsum([len(recs) for _, recs in files]) - 1).tolist()
    made = {}
    for fi, (fn, recs) in enumerate(files):
        iv = T.file_intervals(sel, recs)
        if iv is None:
            continue
        lists = hs.locate(cfg["K"], cfg["E"], revcompl=cfg["rc"], seq_to_file=stf, file_no=fi, intervals=iv)
        cum = np.concatenate([[0], np.cumsum([len(c) for _, c in recs])])
        made[fn + ".genmap.csv"] = T.csv_render(lists, cum, names, last, cfg["rc"])
    gold = os.path.join(folder, "csv")
    assert set(made) == set(os.listdir(gold))
    for fn, text in made.items():
        assert text == open(os.path.join(gold, fn)).read(), (case, fn)


# ---- part lengths / block size chosen by the expected-fetch model (gmb_host.cpp: choose_part_lengths) -----------
def test_part_length_model_changes_the_tree_not_the_counts(monkeypatch):
    import genmap_b200 as gm
    seqs = gm.synth_genome(2_000_000, 2, 77)
    hs = T.HostSim(seqs)
    res = {}
    for name, env in (("equal", {"GMB_PART_MODEL": "0"}), ("model", {}), ("forced", {"GMB_PART_WEIGHTS": "9,1,1,9"})):
        monkeypatch.delenv("GMB_PART_MODEL", raising=False)
        monkeypatch.delenv("GMB_PART_WEIGHTS", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for K, E in ((24, 2), (36, 3), (24, 1)):
            out, f = hs.map(K, E, pos_begin=500_000, pos_end=500_000 + (400 if E == 3 else 4000), return_fetches=True)
            res[name, K, E] = (out, f, f + hs.last_lut_reads + hs.last_fetch_stats[12])  # all memory accesses: + table and text reads
    for K, E in ((24, 2), (36, 3), (24, 1)):
        assert np.array_equal(res["equal", K, E][0], res["model", K, E][0]), (K, E)
        assert np.array_equal(res["equal", K, E][0], res["forced", K, E][0]), (K, E)
    # fewer rank-block fetches with the model where the scheme has more than two parts; at E = 1 never more memory
    # accesses in total (the model weighs rank blocks, table entries and text reads alike)
    assert res["model", 24, 2][1] < 0.85 * res["equal", 24, 2][1]
    assert res["model", 36, 3][1] < 0.9 * res["equal", 36, 3][1]
    assert res["model", 24, 1][2] <= 1.02 * res["equal", 24, 1][2], (res["model", 24, 1][1:], res["equal", 24, 1][1:])


def test_substituted_jump_table_keys_change_the_accesses_not_the_counts(monkeypatch):
    """Entering a search once per admissible string (JumpPlan.variants) against walking from its error-free prefix:
    identical counts, fewer rank-block fetches, more table reads; all E, both value types, -ep, selection."""
    import genmap_b200 as gm
    seqs = gm.synth_genome(2_000_000, 2, 78)
    hs = T.HostSim(seqs, with_sa=True)
    stf = np.array([0, 1], dtype=np.uint32)
    res = {}
    for name, val in (("walk", "0"), ("keys", "1")):
        monkeypatch.setenv("GMB_JUMP_VARIANTS", val)
        for K, E, n in ((24, 1, 4000), (24, 2, 3000), (30, 3, 300), (36, 4, 60)):
            out, f = hs.map(K, E, pos_begin=400_000, pos_end=400_000 + n, return_fetches=True)
            res[name, K, E] = (out, f, hs.last_lut_reads)
        res[name, "ep"] = hs.map(20, 2, seq_to_file=stf, file_no=1, exclude_pseudo=True, pos_begin=1000, pos_end=3000, value_bits=8)
        res[name, "sel"] = hs.map(22, 2, revcompl=False, intervals=[(500, 900), (999_990, 1_000_400)])
    for key in [k for k in res if k[0] == "walk"]:
        a, b = res[key], res[("keys",) + key[1:]]
        if isinstance(a, tuple):
            assert np.array_equal(a[0], b[0]), key
        else:
            assert np.array_equal(a, b), key
    for K, E in ((24, 2), (30, 3), (36, 4)):
        walk, keys = res["walk", K, E], res["keys", K, E]
        assert keys[1] + keys[2] < 0.85 * (walk[1] + walk[2]), (K, E, walk[1:], keys[1:])  # fewer accesses in total
        assert keys[2] > walk[2]                                                            # ... through more table reads


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_state_machine_fuzz_against_the_definition(seed, monkeypatch):
    """Random small genomes (Dna4 and Dna5) x random (K, E, strand, block size, jump depth, value type, table entries
    with / without substituted keys, -ep) against the definition-level counter."""
    rng = np.random.default_rng(seed)
    for _ in range(14):
        nfiles, per, length = int(rng.integers(1, 4)), int(rng.integers(1, 3)), int(rng.integers(30, 600))
        base = T.repeat_rich(int(rng.integers(0, 1 << 30)), per, length, with_n=rng.random() < 0.3)
        seqs, stf = [], []
        for g in range(nfiles):
            for b in base:
                b = b.copy()
                m = rng.random(len(b)) < 0.03 * g
                b[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
                seqs.append(b); stf.append(g)
        stf = np.array(stf, dtype=np.uint32)
        hs = T.HostSim(seqs, with_sa=True)
        for _ in range(4):
            E = int(rng.integers(0, 5)); K = int(rng.integers(E + 2, 36))
            if K * (1 + E) > 80 and length > 300:
                continue
            rc, B = bool(rng.random() < 0.7), int(rng.integers(0, 7))
            depth, bits = int(rng.choice([-1, -1, 1, 3, 8])), int(rng.choice([8, 16]))
            ep = nfiles > 1 and rng.random() < 0.4
            fi = int(rng.integers(0, nfiles))
            monkeypatch.setenv("GMB_JUMP_VARIANTS", "1" if rng.random() < 0.8 else "0")
            want = T.brute(seqs, K, E, revcompl=rc, value_bits=bits, exclude_pseudo=ep, seq_to_file=stf, file_no=fi)
            got = hs.map(K, E, revcompl=rc, value_bits=bits, block_kmers=B, jump_depth=depth, exclude_pseudo=ep, seq_to_file=stf, file_no=fi)
            assert np.array_equal(got, want), dict(seed=seed, K=K, E=E, rc=rc, B=B, depth=depth, bits=bits, ep=ep, fi=fi)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_located_entries_on_fragmented_genomes(seed, monkeypatch):
    """Located table entries on genomes made of many short sequences: candidate alignments, their windows and the
    entries' context characters cross sequence boundaries all the time; the two-phase driver and the general state machine,
    with and without located entries, against the oracle."""
    rng = np.random.default_rng(seed)
    seqs = T.fragmented_genome(seed, 9000)
    orc, hs = T.Oracle(seqs), T.HostSim(seqs)
    for _ in range(6):
        E = int(rng.integers(0, 4)); K = int(rng.integers(max(E + 2, 8), 34))
        rc, B, bits = bool(rng.random() < 0.7), int(rng.integers(0, 6)), int(rng.choice([8, 16]))
        want = orc.map(K, E, revcompl=rc, value_bits=bits)
        for env in ({}, {"GMB_LOCATE": "0"}, {"GMB_BLOCK_KERNEL": "0"}):
            for k in ("GMB_LOCATE", "GMB_BLOCK_KERNEL"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got = hs.map(K, E, revcompl=rc, value_bits=bits, block_kmers=B)
            assert np.array_equal(got, want), dict(seed=seed, K=K, E=E, rc=rc, B=B, bits=bits, env=env, at=np.nonzero(got != want)[0][:8])
        assert hs.last_fetch_stats[11] >= 0
