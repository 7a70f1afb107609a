"""Parity of the CUDA path (through the C ABI) with the oracle, the reference's golden vectors and the
outputs of the unmodified reference binary.  Needs a B200: run with -m gpu."""
import hashlib
import os

import numpy as np
import pytest

import gmtest as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gm():
    import genmap_b200
    from genmap_b200 import _build, _lib
    _build.build()
    assert _lib.lib().gmb_device_count() > 0, "no CUDA device"
    return genmap_b200


def _map(gm, ix, K, E, rc=True, bits=16, limits=None, stf=None, file_no=0, intervals=None, **kw):
    stf_, tb, tl, cum, iv = T._prep(limits, stf, file_no, intervals)
    return ix.compute_mappability(gm.SearchParams(K, E, rc, False, bits), text_begin=tb, text_len=tl,
                                  chrom_cum_lengths=cum, intervals=iv, **kw)


# ---- index builders ---------------------------------------------------------------------------------
@pytest.mark.parametrize("with_n", [False, True], ids=["dna4", "dna5"])
@pytest.mark.parametrize("seed,nchr,length", [(1, 1, 400), (2, 3, 1500), (3, 5, 77), (4, 2, 20000), (6, 4, 250000)])
def test_gpu_index_builder_is_bit_identical_to_host_builder(gm, seed, nchr, length, with_n):
    seqs = T.repeat_rich(seed, nchr, length, with_n=with_n)
    if with_n:
        seqs[0][5:45] = 4  # a run of N as in an assembly gap
    host = gm.Index.build_blob(seqs, with_sa=True)
    dev = gm.Index.build_blob(seqs, with_sa=True, on_gpu=True)
    assert host.nbytes == dev.nbytes
    if host.tobytes() != dev.tobytes():
        diff = np.nonzero(host != dev)[0]
        raise AssertionError("blobs differ at %d bytes, first offsets %s" % (len(diff), diff[:8]))


def test_gpu_index_builder_degenerate_texts(gm):
    for seqs in ([np.zeros(5000, np.uint8)], [np.tile(np.array([0, 1], np.uint8), 3000)] * 3,
                 [np.array([2], np.uint8), np.array([2], np.uint8)], [np.full(193, 3, np.uint8), np.full(191, 3, np.uint8)]):
        assert gm.Index.build_blob(seqs, with_sa=True).tobytes() == gm.Index.build_blob(seqs, with_sa=True, on_gpu=True).tobytes()


# ---- the reference's golden vectors (cases without -ep; 1c-1g are Dna5 genomes) ---------------------
GOLDEN_CASES = ["1a", "1b", "1c", "1d", "1e", "1f", "1g", "2a", "2b", "2c", "2d", "2e", "3a", "3b"]


@pytest.mark.parametrize("case", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [16, 8])
def test_cuda_matches_reference_golden(gm, case, bits):
    cfg = T.CASES[case]
    files, sel, folder = T.load_case(case)
    seqs, stf, _ = T.case_layout(files)
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs, on_gpu=bits == 8)
    ext = "freq16" if bits == 16 else "freq8"
    for fi, (base, recs) in enumerate(files):
        iv = T.file_intervals(sel, recs)
        if iv is None:
            continue
        gold = np.fromfile(os.path.join(folder, "raw_" + ext, base + ".genmap." + ext),
                           dtype=np.uint16 if bits == 16 else np.uint8)
        got = _map(gm, ix, cfg["K"], cfg["E"], rc=cfg["rc"], bits=bits, limits=limits, stf=stf, file_no=fi, intervals=iv)
        assert np.array_equal(got, gold), (case, base)


# ---- outputs of the unmodified reference binary (fixtures) -----------------------------------------
import test_ref_fixtures as RF  # noqa: E402


@pytest.mark.parametrize("line", [c for c in RF.CASES if "-ep" not in c],
                         ids=lambda c: c.split("|")[0])
def test_cuda_matches_reference_binary_fixtures(gm, line):
    name, K, E, flags, bits, seqs, stf, outs = RF.load_fixture(line)
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs, on_gpu=True)
    for fi, gold in enumerate(outs):
        got = _map(gm, ix, K, E, rc="-nc" not in flags, bits=bits, limits=limits, stf=stf, file_no=fi)
        assert np.array_equal(got, gold), (name, fi, np.nonzero(got != gold)[0][:10])


# ---- seeded inputs against the oracle ----------------------------------------------------------------
@pytest.mark.parametrize("K,E", [(30, 0), (30, 1), (30, 2), (21, 3), (16, 4), (50, 2), (65, 1), (130, 3), (200, 2),
                                 (2, 0), (3, 1), (4, 2), (5, 3), (6, 4), (32, 2), (33, 2)])
def test_cuda_matches_oracle(gm, K, E):
    seqs = T.repeat_rich(7, 3, 3000)
    _, limits = T.concat(seqs)
    orc = T.Oracle(seqs)
    ix = gm.Index.build(seqs)
    for rc in (True, False):
        got = _map(gm, ix, K, E, rc=rc, limits=limits)
        want = orc.map(K, E, revcompl=rc)
        assert np.array_equal(got, want), (K, E, rc, np.nonzero(got != want)[0][:10])


@pytest.mark.parametrize("K,E", [(20, 0), (20, 1), (20, 2), (14, 3), (9, 4), (30, 2), (40, 1), (70, 2), (140, 1), (2, 0), (3, 1)])
def test_cuda_dna5_matches_oracle(gm, K, E):
    """Genomes with N (Dna5 index, src/indexing.hpp:459-473): N in the text is an ordinary fifth symbol, N in the
    k-mer never matches (SeqAn's ordValue comparison is on the text side only through the index alphabet)."""
    seqs = T.repeat_rich(11, 3, 2500, with_n=True)
    seqs[1][100:160] = 4
    seqs[2][-5:] = 4
    _, limits = T.concat(seqs)
    orc, ix = T.Oracle(seqs), gm.Index.build(seqs)
    assert ix.info.alphabet_size == 5 and ix.info.rank_block_bytes == 32
    for rc in (True, False):
        want = orc.map(K, E, revcompl=rc)
        for B in (1, 3, 0):
            for depth in (0, -1):
                ix.set_jump_depth(depth)
                p = gm.SearchParams(K, E, rev_compl=rc, block_kmers=B)
                got = ix.compute_mappability(p, chrom_cum_lengths=limits)
                assert np.array_equal(got, want), (K, E, rc, B, depth, np.nonzero(got != want)[0][:10])


def _dna5_genome_with_gaps():
    seqs = T.repeat_rich(51, 3, 2500, with_n=True)
    seqs[1][100:160] = 4
    seqs[1][300:302] = 4
    seqs[2][-5:] = 4
    seqs[0][[7, 500, 501, 1200, 1230]] = 4
    return seqs


@pytest.mark.parametrize("K,E", [(20, 1), (20, 2), (14, 3), (9, 4), (30, 2), (40, 1), (33, 2), (70, 2), (130, 1)])
def test_cuda_dna5_searches_that_skip_the_text_n_plus_the_n_pass(gm, K, E, monkeypatch):
    """Dna5 index WITH the suffix array, E >= 1: the searches never match a text N, so they enter through
    substituted keys and run in the two-phase kernel like on a Dna4 index; the N pass then locates the text windows with
    1..E N through the index and adds the alignments to them.  Gap edges, a short run, isolated N, N at a sequence end,
    both strands, every kernel, 8-bit saturation, selection intervals, position slices — against the oracle."""
    seqs = _dna5_genome_with_gaps()
    _, limits = T.concat(seqs)
    orc, ix = T.Oracle(seqs), gm.Index.build(seqs, with_sa=True)
    try:
        for env in ({}, {"GMB_BLOCK_KERNEL": "0"}):
            monkeypatch.delenv("GMB_BLOCK_KERNEL", raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            for rc in (True, False):
                want = orc.map(K, E, revcompl=rc)
                for B in (1, 3, 0):
                    for depth in (0, 5, -1):
                        ix.set_jump_depth(depth)
                        p = gm.SearchParams(K, E, rev_compl=rc, block_kmers=B)
                        got = ix.compute_mappability(p, chrom_cum_lengths=limits)
                        assert np.array_equal(got, want), (K, E, rc, B, depth, env, np.nonzero(got != want)[0][:10])
        monkeypatch.delenv("GMB_BLOCK_KERNEL", raising=False)
        ix.set_jump_depth(-1)
        iv = np.array([[90, 400], [2400, 2600], [5100, 7400]], dtype=np.uint64)
        p8 = gm.SearchParams(K, E, True, False, 8)
        assert np.array_equal(ix.compute_mappability(p8, chrom_cum_lengths=limits, intervals=iv), orc.map(K, E, value_bits=8, intervals=iv))
        want = orc.map(K, E)
        for a, b in ((0, 97), (97, 1201), (1201, 7500)):
            got = ix.compute_mappability_range(gm.SearchParams(K, E), a, b, chrom_cum_lengths=limits)
            assert np.array_equal(got, want[a:b]), (K, E, a, b)
        start, value = ix.compute_runs(gm.SearchParams(K, E), chrom_cum_lengths=limits)
        ends = np.append(start[1:], len(want)).astype(np.int64)
        assert np.array_equal(np.repeat(value, ends - start.astype(np.int64)), want)
    finally:
        _close(ix)


def test_cuda_dna5_n_pass_on_several_files_and_fragmented_genomes(gm):
    """The N pass is built once for the whole index and applied per FASTA file (text_begin > 0); sequences of a few
    dozen bases with N anywhere (windows and located contexts crossing sequence ends)."""
    base = _dna5_genome_with_gaps()
    rng = np.random.default_rng(9)
    seqs, stf = [], []
    for g in range(3):
        for b in base:
            b = b.copy()
            m = (rng.random(len(b)) < 0.02 * g) & (b < 4)
            b[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
            seqs.append(b); stf.append(g)
    stf = np.array(stf, dtype=np.uint32)
    _, limits = T.concat(seqs)
    orc, ix = T.Oracle(seqs, seq_to_file=stf), gm.Index.build(seqs, with_sa=True, seq_to_file=stf)
    try:
        for K, E in ((24, 2), (30, 1)):
            for f in (0, 1, 2):
                got = _map(gm, ix, K, E, limits=limits, stf=stf, file_no=f)
                assert np.array_equal(got, orc.map(K, E, file_no=f)), (K, E, f)
    finally:
        _close(ix)
    seqs = T.fragmented_genome(4, 14000, with_n=True)
    _, limits = T.concat(seqs)
    orc, ix = T.Oracle(seqs), gm.Index.build(seqs, with_sa=True)
    try:
        for _ in range(12):
            E = int(rng.integers(1, 5)); K = int(rng.integers(max(E + 2, 8), 40))
            if E == 4 and K > 20:
                K = 20
            rc, B, bits = bool(rng.random() < 0.7), int(rng.integers(0, 7)), int(rng.choice([8, 16]))
            ix.set_plan_text_size(int(rng.choice([0, 4 ** 9, 4 ** 12 - 1])))
            want = orc.map(K, E, revcompl=rc, value_bits=bits)
            got = ix.compute_mappability(gm.SearchParams(K, E, rc, False, bits, block_kmers=B), chrom_cum_lengths=limits)
            assert np.array_equal(got, want), dict(K=K, E=E, rc=rc, B=B, bits=bits, at=np.nonzero(got != want)[0][:8])
    finally:
        _close(ix)


def test_cuda_dna5_with_n_everywhere_falls_back_to_walking_the_n_children(gm):
    """More windows with N than the N pass takes (1/64 of the text): the call runs as on an index without the suffix
    array — same counts."""
    seqs = T.repeat_rich(61, 2, 40000, with_n=True)
    for s in seqs:
        s[::37] = 4
    _, limits = T.concat(seqs)
    orc, ix = T.Oracle(seqs), gm.Index.build(seqs, with_sa=True)
    try:
        for K, E in ((20, 1), (24, 2)):
            got, st = ix.compute_mappability(gm.SearchParams(K, E), chrom_cum_lengths=limits, return_stats=True)
            assert np.array_equal(got, orc.map(K, E)), (K, E)
            assert st.kernel_launches == 1  # no N pass
    finally:
        _close(ix)


def test_cuda_dna5_n_pass_at_2mbp_equals_the_walked_n_children_and_the_host_mirror(gm, monkeypatch):
    """The two ways of treating the text's N give the same counts on a 2 Mbp genome with assembly gaps and short runs,
    and the device reads what the host mirror reads."""
    seqs = gm.synth_genome(2_000_000, 4, 77)
    rng = np.random.default_rng(3)
    for s in seqs:
        a = int(rng.integers(0, len(s) - 60000))
        s[a:a + 50000] = 4
        for a in rng.integers(0, len(s) - 100, 12):
            s[int(a):int(a) + int(rng.integers(1, 40))] = 4
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs, with_sa=True)
    try:
        for K, E in ((30, 1), (30, 2), (50, 2)):
            monkeypatch.setenv("GMB_DNA5_NFREE", "0")
            walked, st0 = ix.compute_mappability(gm.SearchParams(K, E), chrom_cum_lengths=limits, count_fetches=True, return_stats=True)
            monkeypatch.delenv("GMB_DNA5_NFREE")
            got, st1 = ix.compute_mappability(gm.SearchParams(K, E), chrom_cum_lengths=limits, count_fetches=True, return_stats=True)
            assert np.array_equal(got, walked), (K, E, np.nonzero(got != walked)[0][:10])
            assert (st0.kernel_launches, st1.kernel_launches) == (1, 3), (K, E)  # search kernel + the two kernels of the N pass
    finally:
        _close(ix)
    small = _dna5_genome_with_gaps()
    _, limits = T.concat(small)
    ix, hs = gm.Index.build(small, with_sa=True), T.HostSim(small, with_sa=True)
    try:
        for K, E, B in ((20, 2, 4), (20, 1, 3), (33, 2, 3), (33, 2, 1), (40, 2, 3)):
            ix.set_jump_depth(5)
            out, st = ix.compute_mappability(gm.SearchParams(K, E, block_kmers=B), chrom_cum_lengths=limits, count_fetches=True, return_stats=True)
            want = hs.map(K, E, jump_depth=5, block_kmers=B)
            assert np.array_equal(out, want)
            # (the two-phase kernel reads an entry again when a round puts more than eight aside: >= for the table reads)
            assert st.rank_block_fetches == hs.last_fetch_stats[0] and st.jump_table_reads >= hs.last_lut_reads, (K, E, B)
    finally:
        _close(ix)


def test_cuda_dna5_exclude_pseudo_bwt_export_and_fetch_counter(gm):
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
    _, limits = T.concat(ms)
    mo, ix, hs = T.Oracle(ms, seq_to_file=stf), gm.Index.build(ms, with_sa=True, seq_to_file=stf), T.HostSim(ms, with_sa=True)
    for rev in (False, True):
        assert np.array_equal(ix.export_bwt(rev), mo.bwt(rev))
    assert np.array_equal(ix.export_sa(), mo.sa().astype(np.uint32))
    for f in (0, 2):
        want = mo.map(22, 2, exclude_pseudo=True, file_no=f)
        stf_, tb, tl, cum, _ = T._prep(limits, stf, f, None)
        for B in (1, 4):
            got = ix.compute_mappability(gm.SearchParams(22, 2, True, True, 16, block_kmers=B), text_begin=tb, text_len=tl,
                                         chrom_cum_lengths=cum)
            assert np.array_equal(got, want), (f, B)
    for K, E in [(20, 0), (20, 2)]:
        out, st = _map(gm, ix, K, E, limits=limits, count_fetches=True, return_stats=True)
        want, f = hs.map(K, E, return_fetches=True)
        # (the index holds the suffix array: at E = 2 both sides skip the text's N and run the two-phase driver, whose
        # device form reads an entry again when a round puts more than eight aside: >= for the table reads)
        assert np.array_equal(out, want) and st.rank_block_fetches == f and st.jump_table_reads >= hs.last_lut_reads


def test_cuda_dna5_matches_reference_binary_live(gm, tmp_path):
    if not T.have_reference():
        pytest.skip("oracle/_ref/genmap_ref not present")
    seqs = gm.synth_genome(600_000, 3, 123)
    rng = np.random.default_rng(8)
    for s in seqs:  # gaps and scattered N
        a = int(rng.integers(0, len(s) - 5000))
        s[a:a + 3000] = 4
        s[rng.integers(0, len(s), 40)] = 4
    fa = str(tmp_path / "g.fa")
    T.write_fasta(fa, seqs)
    ix = gm.Index.build(seqs)
    assert ix.info.alphabet_size == 5
    for K, E in [(30, 0), (24, 1), (36, 2)]:
        ref = T.run_reference(fa, K, E)["g"]
        got = ix.compute_mappability(gm.SearchParams(K, E))
        assert np.array_equal(got, ref), (K, E)


def test_cuda_jump_table_depths_do_not_change_results(gm):
    seqs = T.repeat_rich(31, 3, 6000)
    _, limits = T.concat(seqs)
    ix, orc = gm.Index.build(seqs), T.Oracle(seqs)
    for K, E in [(30, 0), (30, 2), (16, 3), (50, 2), (13, 0)]:
        for rc in (True, False):
            want = orc.map(K, E, revcompl=rc)
            for depth in (0, 1, 4, 8, 12, -1):
                ix.set_jump_depth(depth)
                assert np.array_equal(_map(gm, ix, K, E, rc=rc, limits=limits), want), (K, E, rc, depth)


@pytest.mark.parametrize("K,E", [(30, 1), (30, 2), (21, 3), (16, 4), (50, 2), (9, 1), (65, 1)])
def test_cuda_block_size_does_not_change_results(gm, K, E):
    """k-mers per block (the reference's -xo / overlap): every value gives the same counts (tests/tests.sh:47-60)."""
    seqs = T.repeat_rich(7, 3, 2500) + [np.array([0, 1, 2], np.uint8), T.repeat_rich(9, 1, K + 3)[0]]
    _, limits = T.concat(seqs)
    orc, ix = T.Oracle(seqs), gm.Index.build(seqs)
    want, want_nc = orc.map(K, E), orc.map(K, E, revcompl=False)
    for B in (1, 2, 3, 5, 8, 16, 0):
        p = gm.SearchParams(K, E, block_kmers=B)
        assert np.array_equal(ix.compute_mappability(p, chrom_cum_lengths=limits), want), (K, E, B)
        p = gm.SearchParams(K, E, rev_compl=False, block_kmers=B)
        assert np.array_equal(ix.compute_mappability(p, chrom_cum_lengths=limits), want_nc), (K, E, B)


def test_cuda_edge_cases(gm):
    # sequences shorter than K between longer ones, selection intervals, saturation, palindromes
    seqs = [np.array([0, 1, 2], np.uint8), T.repeat_rich(3, 1, 300)[0], np.array([3, 3], np.uint8),
            T.repeat_rich(4, 1, 200)[0], np.array([2], np.uint8)]
    _, limits = T.concat(seqs)
    orc, ix = T.Oracle(seqs), gm.Index.build(seqs)
    assert np.array_equal(_map(gm, ix, 12, 1, limits=limits), orc.map(12, 1))
    iv = [(0, 2), (10, 40), (35, 60), (290, 320), (500, 506)]
    assert np.array_equal(_map(gm, ix, 12, 1, limits=limits, intervals=iv), orc.map(12, 1, intervals=iv))
    # K longer than all but one sequence; K longer than every sequence: all zeros
    assert np.array_equal(_map(gm, ix, 250, 0, limits=limits), orc.map(250, 0))
    short = [T.repeat_rich(3, 1, 100)[0], T.repeat_rich(4, 1, 60)[0]]
    _, sl = T.concat(short)
    assert not _map(gm, gm.Index.build(short), 101, 0, limits=sl).any()
    pal = [np.tile(np.array([0, 1, 2, 3], dtype=np.uint8), 2000)]
    _, pl = T.concat(pal)
    po, pix = T.Oracle(pal), gm.Index.build(pal)
    for bits in (8, 16):
        got = _map(gm, pix, 8, 0, bits=bits, limits=pl)
        assert np.array_equal(got, po.map(8, 0, value_bits=bits))
        assert got.max() == (255 if bits == 8 else 3998)
    with pytest.raises(gm.GenmapError):
        _map(gm, ix, 30, 5, limits=limits)  # E > 4 (src/mappability.hpp:187)
    with pytest.raises(gm.GenmapError):
        _map(gm, ix, 3, 2, limits=limits)   # K < E + 2


def test_cuda_fetch_counter_matches_cpu_restatement(gm):
    """The roofline's algorithmic unit: rank-block fetches counted by the instrumented kernel must
    equal the count of the host-compiled state machine on the same input."""
    seqs = T.repeat_rich(21, 2, 40000)
    _, limits = T.concat(seqs)
    ix, hs = gm.Index.build(seqs), T.HostSim(seqs)
    for K, E in [(30, 0), (30, 1), (30, 2)]:
        out, st = _map(gm, ix, K, E, limits=limits, count_fetches=True, return_stats=True)
        want, f = hs.map(K, E, return_fetches=True)
        assert np.array_equal(out, want)
        assert st.rank_block_fetches == f, (K, E, st.rank_block_fetches, f)
        assert st.jump_table_reads == hs.last_lut_reads
    for depth in (0, 3, 9):
        ix.set_jump_depth(depth)
        out, st = _map(gm, ix, 30, 1, limits=limits, count_fetches=True, return_stats=True)
        want, f = hs.map(30, 1, return_fetches=True, jump_depth=depth)
        assert np.array_equal(out, want) and st.rank_block_fetches == f and st.jump_depth == min(depth, 9)


@pytest.mark.parametrize("with_n", [False, True], ids=["dna4", "dna5"])
def test_cuda_located_entries_are_used_like_in_the_host_mirror(gm, with_n, monkeypatch):
    """Located table entries on Dna4 and Dna5 indices, general kernel: the device verifies exactly the
    entries the host mirror verifies (same rank-block fetches, table reads, located entries, text reads) — a table
    whose text pass located nothing would still give the right counts, only slower."""
    monkeypatch.setenv("GMB_BLOCK_KERNEL", "0")  # the two-phase kernel keeps its own counters
    seqs = T.repeat_rich(23, 2, 60000, rep_frac=0.05, with_n=with_n)
    _, limits = T.concat(seqs)
    ix, hs = gm.Index.build(seqs), T.HostSim(seqs)
    try:
        bad = []
        for K, E, B, depth in [(30, 1, 1, -1), (30, 1, 3, -1), (30, 1, 3, 7), (24, 2, 4, -1), (30, 0, 1, -1)]:
            ix.set_jump_depth(depth)
            p = gm.SearchParams(K, E, block_kmers=B)
            out, st = ix.compute_mappability(p, chrom_cum_lengths=limits, count_fetches=True, return_stats=True)
            want = hs.map(K, E, jump_depth=depth, block_kmers=B)
            f = hs.last_fetch_stats
            got = (st.rank_block_fetches, st.jump_table_reads, st.located_entries, st.text_reads)
            host = (f[0], hs.last_lut_reads, f[11], f[12])
            if not np.array_equal(out, want) or got != host or st.located_entries == 0:
                bad.append(dict(K=K, E=E, B=B, depth=depth, equal=bool(np.array_equal(out, want)), device=got, host=host))
        assert not bad, bad
    finally:
        _close(ix)


# ---- 1 Mbp of the frozen synthetic generator: md5 pins measured with the reference --
PINS = {0: "15a50bb1184e42daacd5569323356621", 1: "27e5f62a996ee570ba5e600972303903", 2: "4445ff36c72e8ff1a45f3ab08685d0c4"}


@pytest.mark.parametrize("E", [0, 1, 2])
def test_cuda_1mbp_matches_reference_md5(gm, E):
    seqs = gm.synth_genome(1_000_000, 1, 42)
    ix = gm.Index.build(seqs)
    got = ix.compute_mappability(gm.SearchParams(30, E))
    assert hashlib.md5(got.tobytes()).hexdigest() == PINS[E]


def test_cuda_matches_reference_binary_live(gm, tmp_path):
    """Run the unmodified reference binary on the box (it travels in oracle/_ref) and compare."""
    if not T.have_reference():
        pytest.skip("oracle/_ref/genmap_ref not present")
    seqs = gm.synth_genome(2_000_000, 3, 99)
    fa = str(tmp_path / "g.fa")
    T.write_fasta(fa, seqs)
    ix = gm.Index.build(seqs)
    for K, E in [(30, 0), (24, 1), (36, 2)]:
        ref = T.run_reference(fa, K, E)["g"]
        got = ix.compute_mappability(gm.SearchParams(K, E))
        assert np.array_equal(got, ref), (K, E)


def test_cuda_sharded_ranges_compose(gm):
    import torch
    seqs = gm.synth_genome(400_000, 4, 5)
    ix = gm.Index.build(seqs)
    p = gm.SearchParams(30, 1)
    whole = ix.compute_mappability(p)
    buf = torch.zeros(ix.n_text, dtype=torch.int16, device="cuda:0")
    cuts = [0, 1, 99_999, 100_010, 250_000, ix.n_text]
    for b, e in zip(cuts[:-1], cuts[1:]):
        ix.compute_mappability_device(p, buf.data_ptr(), pos_begin=b, pos_end=e,
                                      stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy().view(np.uint16), whole)


def test_cuda_range_slices_and_bwt_export(gm):
    seqs = T.repeat_rich(9, 3, 5000)
    ix, orc = gm.Index.build(seqs), T.Oracle(seqs)
    p = gm.SearchParams(20, 1)
    whole = ix.compute_mappability(p)
    for b, e in [(0, 10), (4990, 5020), (7000, 15000)]:
        assert np.array_equal(ix.compute_mappability_range(p, b, e), whole[b:e])
    for rev in (False, True):
        assert np.array_equal(ix.export_bwt(rev), orc.bwt(rev))


# ---- --exclude-pseudo (locate + distinct FASTA files) -------------------------------------------------
@pytest.mark.parametrize("case", ["3c", "3d", "3e", "3f"])
def test_cuda_exclude_pseudo_matches_reference_golden(gm, case):
    cfg = T.CASES[case]
    files, sel, folder = T.load_case(case)
    seqs, stf, _ = T.case_layout(files)
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs, with_sa=True, seq_to_file=stf)
    for fi, (base, recs) in enumerate(files):
        iv = T.file_intervals(sel, recs)
        if iv is None:
            continue
        gold = np.fromfile(os.path.join(folder, "raw_freq16", base + ".genmap.freq16"), dtype=np.uint16)
        stf_, tb, tl, cum, ivv = T._prep(limits, stf, fi, iv)
        got = ix.compute_mappability(gm.SearchParams(cfg["K"], cfg["E"], cfg["rc"], True, 16), text_begin=tb, text_len=tl,
                                     chrom_cum_lengths=cum, intervals=ivv)
        assert np.array_equal(got, gold), (case, base)


@pytest.mark.parametrize("line", [c for c in RF.CASES if "-ep" in c], ids=lambda c: c.split("|")[0])
def test_cuda_exclude_pseudo_matches_reference_binary_fixtures(gm, line):
    name, K, E, flags, bits, seqs, stf, outs = RF.load_fixture(line)
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs, with_sa=True, seq_to_file=stf)
    for fi, gold in enumerate(outs):
        stf_, tb, tl, cum, _ = T._prep(limits, stf, fi, None)
        got = ix.compute_mappability(gm.SearchParams(K, E, "-nc" not in flags, True, bits), text_begin=tb, text_len=tl,
                                     chrom_cum_lengths=cum)
        assert np.array_equal(got, gold), (name, fi)


def test_cuda_exclude_pseudo_pangenome_scale_model(gm):
    """BASELINE config 5 in miniature: 10 FASTA files x 3 chromosomes, file g = base with g % substitutions,
    K=50 E=2 --exclude-pseudo, against the oracle."""
    base = T.repeat_rich(46, 3, 3000)
    rng = np.random.default_rng(46)
    seqs, stf = [], []
    for g in range(10):
        for s in base:
            s = s.copy()
            m = rng.random(len(s)) < 0.01 * g
            s[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
            seqs.append(s); stf.append(g)
    stf = np.array(stf, dtype=np.uint32)
    _, limits = T.concat(seqs)
    orc = T.Oracle(seqs, seq_to_file=stf)
    ix = gm.Index.build(seqs, with_sa=True, seq_to_file=stf)
    for fi in (0, 4, 9):
        stf_, tb, tl, cum, _ = T._prep(limits, stf, fi, None)
        got = ix.compute_mappability(gm.SearchParams(50, 2, True, True, 16), text_begin=tb, text_len=tl, chrom_cum_lengths=cum)
        assert np.array_equal(got, orc.map(50, 2, exclude_pseudo=True, file_no=fi)), fi
    with pytest.raises(gm.GenmapError):  # no suffix array in the index
        gm.Index.build(seqs, with_sa=False, seq_to_file=stf).compute_mappability(gm.SearchParams(50, 2, True, True, 16), text_begin=0,
                                                                                  text_len=int(limits[3]), chrom_cum_lengths=limits[:4])


def test_gpu_built_index_feeds_the_unmodified_reference(gm, tmp_path):
    """BWT + SA from the GPU builder, written in the reference's on-disk format, make the unmodified
    reference binary produce the same frequencies as the CUDA path (this is what bench.py's CPU arm does)."""
    if not T.have_reference():
        pytest.skip("oracle/_ref/genmap_ref not present")
    import subprocess
    seqs = gm.synth_genome(1_200_000, 3, 7)
    ix = gm.Index.build(seqs, with_sa=True)
    files = [("genome.fa", [("chr%d" % (i + 1), s) for i, s in enumerate(seqs)])]
    d = T.write_seqan_index(str(tmp_path / "index"), files, ix.export_bwt(False), ix.export_bwt(True), ix.export_sa())
    out = tmp_path / "out"
    out.mkdir()
    bed = tmp_path / "w.bed"
    bed.write_text("chr2\t1000\t151000\n")
    subprocess.run([T.REF_BIN, "map", "-I", d, "-O", str(out), "-K", "30", "-E", "1", "-r", "-fl", "-S", str(bed)], check=True,
                   stdout=subprocess.DEVNULL)
    ref = np.fromfile(str(out / "genome.genmap.freq16"), dtype=np.uint16)
    got = ix.compute_mappability(gm.SearchParams(30, 1), intervals=[(401000, 551000)])
    assert np.array_equal(got, ref)


# ---- locations (csv lists): gmb_map_locations ------------------------------------------------------------------
@pytest.mark.parametrize("with_n", [False, True], ids=["dna4", "dna5"])
@pytest.mark.parametrize("K,E,rc", [(12, 0, True), (14, 2, True), (10, 2, False), (33, 3, True), (70, 1, True)])
def test_cuda_locations_match_definition(gm, K, E, rc, with_n):
    seqs = T.repeat_rich(7, 3, 700, with_n=with_n)
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs, with_sa=True)
    off, loc = ix.compute_locations(gm.SearchParams(K, E, rc))
    got = T.lists_from_arrays(off, loc)
    want = T.brute_locations(seqs, K, E, T.valid_starts(limits, K), revcompl=rc)
    assert len(loc) > int(limits[-1]) // 2
    for j in range(int(limits[-1])):
        assert got[j] == want.get(j, ([], [])), j
    # a small budget forces many continuation calls; a range in the middle; an index built on the host
    off2, loc2 = ix.compute_locations(gm.SearchParams(K, E, rc), max_locations=50)
    assert np.array_equal(off, off2) and np.array_equal(loc, loc2)
    off3, loc3 = ix.compute_locations(gm.SearchParams(K, E, rc), pos_begin=300, pos_end=900)
    assert T.lists_from_arrays(off3, loc3, 300) == {j: got[j] for j in range(300, 900)}
    ixh = gm.Index.build(seqs, with_sa=True, on_gpu=False)
    off4, loc4 = ixh.compute_locations(gm.SearchParams(K, E, rc))
    assert np.array_equal(off, off4) and np.array_equal(loc, loc4)


def test_cuda_locations_match_host_state_machine_and_counts(gm):
    """Larger genome: the lists equal the host-compiled state machine's, and their lengths are the unsaturated
    frequencies (countOccurrences summed over itAll, src/algo.hpp:318-326)."""
    seqs = gm.synth_genome(300_000, 3, 11)
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs, with_sa=True)
    hs = T.HostSim(seqs, with_sa=True)
    for K, E in ((30, 0), (30, 2), (20, 1)):
        off, loc = ix.compute_locations(gm.SearchParams(K, E), pos_begin=100_000, pos_end=140_000)
        want = hs.locate(K, E, pos_begin=100_000, pos_end=140_000)
        assert T.lists_from_arrays(off, loc, 100_000) == want, (K, E)
        freq = ix.compute_mappability(gm.SearchParams(K, E))[100_000:140_000]
        n = (off[2::2] - off[:-2:2]).astype(np.int64)
        assert np.array_equal(np.minimum(n, 65535), freq.astype(np.int64)), (K, E)


def test_cuda_locations_need_the_suffix_array(gm):
    ix = gm.Index.build(T.repeat_rich(7, 1, 500), with_sa=False)
    with pytest.raises(gm.GenmapError) as e:
        ix.compute_locations(gm.SearchParams(12, 0))
    assert "suffix array" in str(e.value)


# ---- runs (device run-length encoding for the track writers): gmb_map_runs -------------------------------------
def _host_runs(c, cum, b, e):
    cum = np.asarray(cum, dtype=np.int64)
    head = np.ones(e - b, dtype=bool)
    head[1:] = c[b + 1:e] != c[b:e - 1]
    inside = cum[(cum > b) & (cum < e)] - b
    head[inside] = True
    start = np.nonzero(head)[0] + b
    return start.astype(np.uint64), c[start].astype(np.uint16)


@pytest.mark.parametrize("bits", [16, 8])
def test_cuda_runs_equal_host_scan_of_the_vector(gm, bits):
    seqs = T.repeat_rich(9, 5, 3000) + [np.array([0, 1, 2], np.uint8)] + T.repeat_rich(10, 2, 40)
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs)
    n = int(limits[-1])
    for K, E, iv in ((12, 0, None), (16, 1, None), (10, 2, [(50, 4000), (9000, 9100)])):
        p = gm.SearchParams(K, E, value_bits=bits)
        c = ix.compute_mappability(p, intervals=iv)
        for b, e in ((0, n), (777, 9001), (3000, 3001)):
            start, value = ix.compute_runs(p, pos_begin=b, pos_end=e, intervals=iv)
            ws, wv = _host_runs(c, limits, b, e)
            assert np.array_equal(start, ws) and np.array_equal(value, wv), (K, E, b, e)


def test_cuda_runs_on_a_larger_genome_roundtrip(gm):
    """size-independent property: expanding the runs gives back the vector"""
    seqs = gm.synth_genome(3_000_000, 4, 17)
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs)
    p = gm.SearchParams(30, 1)
    c = ix.compute_mappability(p)
    start, value = ix.compute_runs(p)
    assert len(start) < len(c) // 4
    ends = np.append(start[1:], len(c)).astype(np.int64)
    assert np.array_equal(np.repeat(value, ends - start.astype(np.int64)), c)
    assert np.isin(np.asarray(limits[:-1], dtype=np.uint64), start).all()


def test_cuda_exclude_pseudo_with_more_than_64_files(gm):
    """Beyond the kernel's 64-bit file mask the library locates every occurrence and counts distinct files on the
    device (ep_many_files): 70 single-sequence files + one with three sequences, against the oracle."""
    rng = np.random.default_rng(70)
    base = T.repeat_rich(70, 1, 600)[0]
    seqs, stf = [], []
    for g in range(70):
        s = base.copy()
        m = rng.random(len(s)) < 0.0008 * g
        s[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
        seqs.append(s); stf.append(g)
    for extra in T.repeat_rich(71, 3, 300):
        seqs.append(extra); stf.append(70)
    stf = np.array(stf, dtype=np.uint32)
    _, limits = T.concat(seqs)
    orc = T.Oracle(seqs, seq_to_file=stf)
    ix = gm.Index.build(seqs, with_sa=True, seq_to_file=stf)
    for K, E, rc, bits in ((20, 0, True, 16), (24, 1, True, 8), (16, 2, False, 16)):
        for fi in (0, 33, 69, 70):
            _, tb, tl, cum, _ = T._prep(limits, stf, fi, None)
            got = ix.compute_mappability(gm.SearchParams(K, E, rc, True, bits), text_begin=tb, text_len=tl, chrom_cum_lengths=cum)
            want = orc.map(K, E, revcompl=rc, exclude_pseudo=True, value_bits=bits, file_no=fi)
            assert np.array_equal(got, want), (K, E, fi)
            assert fi != 0 or got.max() > 50


def test_cuda_index_replica_on_a_second_gpu(gm):
    from genmap_b200 import _lib
    if _lib.lib().gmb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    seqs = gm.synth_genome(500_000, 3, 23)
    ix0 = gm.Index.build(seqs, device=0)
    ix1 = ix0.replicate(1)
    assert ix1.device == 1 and int(ix1.info.blob_bytes) == int(ix0.info.blob_bytes)
    for K, E in ((30, 0), (24, 2)):
        assert np.array_equal(ix0.compute_mappability(gm.SearchParams(K, E)), ix1.compute_mappability(gm.SearchParams(K, E)))
    same = ix0.replicate(0)  # a second copy on the same device also works
    assert np.array_equal(same.compute_mappability(gm.SearchParams(30, 1)), ix0.compute_mappability(gm.SearchParams(30, 1)))


# ---- the code paths that only exist at the benchmarked scale, exercised on small genomes -------------------------
# (the 3 Gbp bench runs jump tables of depth 13-16 — 4^16 entries, 16-byte entries with both intervals,
# substituted keys at depth 15/16 — and model-chosen part lengths / block sizes that no small genome selects by
# itself.  The plan depends on the text size only through the planner, so the planner is told a size.)
def _close(ix):
    import torch
    ix.close()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("with_n", [False, True], ids=["dna4", "dna5"])
def test_cuda_plan_of_a_3gbp_genome_on_a_small_one(gm, with_n):
    """The exact search plan of the 3 Gbp bench genome (block sizes, part lengths, entry depths up to 16 with
    substituted keys, 34-69 GB tables) executed on a 3 x 6 kbp genome, against the oracle."""
    seqs = T.repeat_rich(31, 3, 6000, with_n=with_n)
    _, limits = T.concat(seqs)
    ix, orc = gm.Index.build(seqs), T.Oracle(seqs)
    ix.set_plan_text_size(3_000_000_024)
    deepest = 0
    try:
        for K, E in [(30, 0), (30, 1), (30, 2), (50, 2), (36, 3)]:
            for rc in (True, False):
                want = orc.map(K, E, revcompl=rc)
                for B in (0, 1):
                    p = gm.SearchParams(K, E, rev_compl=rc, block_kmers=B)
                    got, st = ix.compute_mappability(p, chrom_cum_lengths=limits, return_stats=True)
                    assert np.array_equal(got, want), (K, E, rc, B, st.jump_depth, np.nonzero(got != want)[0][:10])
                    deepest = max(deepest, int(st.jump_depth))
        assert deepest == 16, deepest
        assert int(ix.refresh_info().jump_table_bytes) > (30 << 30)
    finally:
        _close(ix)


@pytest.mark.parametrize("depth", [13, 14, 15, 16])
def test_cuda_jump_tables_of_depth_13_to_16(gm, depth):
    """Every table flavour (8-byte entries, + SA(T) array, 16-byte entries) at the depths the small-genome tests never
    reach (tables have 4^d entries whatever the genome), E = 0, 1, 2, both alphabets."""
    for with_n in (False, True):
        seqs = T.repeat_rich(33 + depth, 2, 5000, with_n=with_n)
        _, limits = T.concat(seqs)
        ix, orc = gm.Index.build(seqs), T.Oracle(seqs)
        ix.set_plan_text_size(4 ** depth - 1)  # the planner's idea of the text: picks entry depths up to `depth`
        try:
            for K, E in [(30, 0), (30, 1), (30, 2), (24, 2)]:
                want = orc.map(K, E)
                for B in (0, 1):
                    got, st = ix.compute_mappability(gm.SearchParams(K, E, block_kmers=B), chrom_cum_lengths=limits, return_stats=True)
                    assert np.array_equal(got, want), (depth, with_n, K, E, B, st.jump_depth)
                    assert E != 0 or st.jump_depth == depth, (depth, st.jump_depth)
            # a fixed depth instead of a planned one (E = 0 enters at exactly that depth)
            ix.set_plan_text_size(0)
            ix.set_jump_depth(depth)
            got, st = ix.compute_mappability(gm.SearchParams(30, 0), chrom_cum_lengths=limits, return_stats=True)
            assert np.array_equal(got, orc.map(30, 0)) and st.jump_depth == depth
        finally:
            _close(ix)


def test_cuda_plan_cache_and_progress(gm):
    """Cached plans (tables stay on the device between calls) give the same counts when configurations alternate,
    and gmb_progress reports the finished call."""
    seqs = gm.synth_genome(600_000, 3, 77)
    ix = gm.Index.build(seqs)
    assert ix.progress() == (0, 0)
    first = {}
    for rnd in range(3):
        for K, E, bits in [(30, 0, 16), (30, 1, 16), (24, 2, 8), (30, 1, 8)]:
            got = ix.compute_mappability(gm.SearchParams(K, E, value_bits=bits))
            key = (K, E, bits)
            if rnd == 0:
                first[key] = got
            else:
                assert np.array_equal(got, first[key]), key
            done, total = ix.progress()
            assert done == total and 0 < total <= ix.n_text
    assert np.array_equal(np.minimum(first[(30, 1, 16)], 255).astype(np.uint8), first[(30, 1, 8)])
    _close(ix)


def test_cuda_250mbp_whole_file_is_bit_exact_against_the_reference(gm):
    """BASELINE config 2 as BASELINE.md §3 prescribes it: 250 Mbp, K = 30, E = 0, `cmp` of the whole .freq16 against
    the unmodified reference binary; plus windows at E = 1 and E = 2 on the same index.  The reference reads an
    index in its own format written from the GPU builder's BWT + SA (byte-identical to its own `index` output:
    tests/test_seqan_index_writer.py, profiles/r02/writer_identity_40mbp.txt)."""
    if not T.have_reference():
        pytest.skip("oracle/_ref/genmap_ref not present")
    import shutil, subprocess, tempfile
    seqs = gm.synth_genome(250_000_000, 4, 46)
    per = len(seqs[0])
    ix = gm.Index.build(seqs, with_sa=True)
    tmp = tempfile.mkdtemp(prefix="gmb_250_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        files = [("genome.fa", [("chr%d" % (i + 1), s) for i, s in enumerate(seqs)])]
        d = T.write_seqan_index(os.path.join(tmp, "index"), files, ix.export_bwt(False), ix.export_bwt(True), ix.export_sa())
        out = os.path.join(tmp, "out")
        os.mkdir(out)
        base = [T.REF_BIN, "map", "-I", d, "-O", out, "-K", "30", "-r", "-fl", "-T", str(os.cpu_count() or 1)]
        subprocess.run(base + ["-E", "0"], check=True, stdout=subprocess.DEVNULL)
        ref = np.fromfile(os.path.join(out, "genome.genmap.freq16"), dtype=np.uint16)
        got = ix.compute_mappability(gm.SearchParams(30, 0))
        assert len(ref) == len(got) == 4 * per
        assert np.array_equal(got, ref), np.nonzero(got != ref)[0][:10]
        assert int((got > 1).sum()) > 1_000_000  # the repeats are there
        for E, n in ((1, 1_000_000), (2, 100_000)):
            b = per + per // 3
            with open(os.path.join(tmp, "w.bed"), "w") as f:
                f.write("chr2\t%d\t%d\n" % (b - per, b - per + n))
            subprocess.run(base + ["-E", str(E), "-S", os.path.join(tmp, "w.bed")], check=True, stdout=subprocess.DEVNULL)
            ref = np.memmap(os.path.join(out, "genome.genmap.freq16"), dtype=np.uint16, mode="r")
            got = ix.compute_mappability_range(gm.SearchParams(30, E), b, b + n)
            assert np.array_equal(got, ref[b:b + n]), (E, np.nonzero(got != ref[b:b + n])[0][:10])
            del ref
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        _close(ix)


@pytest.mark.parametrize("with_sa", [False, True], ids=["counts", "exclude_pseudo"])
def test_cuda_three_kernels_and_table_flavours_agree(gm, monkeypatch, with_sa):
    """The straight-line E = 0 kernel, the two-phase kernel and the general kernel, with and without located table
    entries, give the same vector (and the oracle's) — also under --exclude-pseudo, with uint8 counts and -nc."""
    seqs = gm.synth_genome(300_000, 3, 31)
    seqs[1][5000:5600] = seqs[0][1000:1600]                      # a planted exact repeat
    seqs[2][100:900] = (3 - seqs[0][2000:2800])[::-1]            # and a reverse-complemented one
    stf = np.array([0, 1, 1], dtype=np.uint32) if with_sa else None
    _, limits = T.concat(seqs)
    ix = gm.Index.build(seqs, with_sa=with_sa, seq_to_file=stf)
    ix.set_plan_text_size(3_000_000_024)  # deep tables, substituted keys: the plan of the bench genome
    orc = T.Oracle(seqs, seq_to_file=stf)
    knobs = [{}, {"GMB_LOCATE": "0"}, {"GMB_BLOCK_KERNEL": "0", "GMB_EXACT_KERNEL": "0"}, {"GMB_BLOCK_KERNEL": "0", "GMB_EXACT_KERNEL": "0", "GMB_LOCATE": "0"},
             {"GMB_BLOCK_KERNEL": "2"}]
    try:
        for K, E, rc, bits in [(30, 0, True, 16), (30, 1, True, 16), (30, 2, True, 8), (24, 3, False, 16), (31, 1, False, 16), (40, 2, True, 16)]:
            fi = 1 if with_sa else 0
            stf_, tb, tl, cum, _ = T._prep(limits, stf, fi, None)
            want = orc.map(K, E, revcompl=rc, exclude_pseudo=with_sa, value_bits=bits, file_no=fi)
            for env in knobs:
                for k in ("GMB_LOCATE", "GMB_BLOCK_KERNEL", "GMB_EXACT_KERNEL"):
                    monkeypatch.delenv(k, raising=False)
                for k, v in env.items():
                    monkeypatch.setenv(k, v)
                got = ix.compute_mappability(gm.SearchParams(K, E, rc, with_sa, bits), text_begin=tb, text_len=tl, chrom_cum_lengths=cum)
                assert np.array_equal(got, want), (K, E, rc, bits, env, np.nonzero(got != want)[0][:10])
    finally:
        _close(ix)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_cuda_located_entries_on_fragmented_genomes(gm, seed):
    """The same on the device, through all three kernels: genomes of many short sequences, random (K, E, strand, block
    size, value type, planner text size), against the oracle."""
    rng = np.random.default_rng(100 + seed)
    with_n = seed == 4
    seqs = T.fragmented_genome(seed, 14000, with_n=with_n)
    _, limits = T.concat(seqs)
    orc, ix = T.Oracle(seqs), gm.Index.build(seqs)
    try:
        for _ in range(10):
            E = int(rng.integers(0, 5)); K = int(rng.integers(max(E + 2, 8), 40))
            if E == 4 and K > 20:
                K = 20
            rc, B, bits = bool(rng.random() < 0.7), int(rng.integers(0, 7)), int(rng.choice([8, 16]))
            ix.set_plan_text_size(int(rng.choice([0, 4 ** 9, 4 ** 12 - 1])))
            want = orc.map(K, E, revcompl=rc, value_bits=bits)
            got = ix.compute_mappability(gm.SearchParams(K, E, rc, False, bits, block_kmers=B), chrom_cum_lengths=limits)
            assert np.array_equal(got, want), dict(seed=seed, K=K, E=E, rc=rc, B=B, bits=bits, at=np.nonzero(got != want)[0][:8])
    finally:
        _close(ix)


def test_cuda_jump_depth_shrinks_when_hbm_is_short(gm):
    """The automatic table depth gives way when HBM is short (the shrink path had no test): with all but
    ~12 GB of the device taken, the plan of a 3 Gbp genome (69 GB of tables at depth 16) is entered at a shallower depth
    and the counts do not change; a cached deeper table is dropped first when another configuration needs the room."""
    import torch
    seqs = T.repeat_rich(41, 3, 5000)
    _, limits = T.concat(seqs)
    ix, orc = gm.Index.build(seqs), T.Oracle(seqs)
    ix.set_plan_text_size(3_000_000_024)
    hog = None
    try:
        got, st = ix.compute_mappability(gm.SearchParams(30, 0), chrom_cum_lengths=limits, return_stats=True)
        assert st.jump_depth == 16 and np.array_equal(got, orc.map(30, 0))
        torch.cuda.synchronize()
        free, _total = torch.cuda.mem_get_info()
        hog = torch.empty(max(0, free - (12 << 30)), dtype=torch.uint8, device="cuda")  # the depth-16 table stays, 12 GB are left
        for K, E in ((30, 1), (30, 2), (50, 2)):
            got, st = ix.compute_mappability(gm.SearchParams(K, E), chrom_cum_lengths=limits, return_stats=True)
            assert np.array_equal(got, orc.map(K, E)), (K, E, st.jump_depth)
            assert 1 <= st.jump_depth <= 16
        del hog
        hog = None
        torch.cuda.empty_cache()
        free, _total = torch.cuda.mem_get_info()
        hog = torch.empty(max(0, free - (6 << 30)), dtype=torch.uint8, device="cuda")
        ix2 = gm.Index.build(T.repeat_rich(42, 2, 4000))
        ix2.set_plan_text_size(3_000_000_024)
        got, st = ix2.compute_mappability(gm.SearchParams(24, 1), return_stats=True)
        assert st.jump_depth < 16 and np.array_equal(got, T.Oracle(T.repeat_rich(42, 2, 4000)).map(24, 1)), st.jump_depth
        ix2.close()
    finally:
        del hog
        _close(ix)
