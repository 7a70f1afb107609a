"""The reference-format index writer (oracle/seqan_index.c + gmtest.write_seqan_index) must reproduce the
files written by the unmodified reference's own `genmap index` byte for byte, and the reference's `map`
must give the same frequencies on it.  Needs oracle/_ref/genmap_ref (present in the build container and,
because oracle/_ref travels, on the GPU box)."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

import gmtest as T

pytestmark = pytest.mark.skipif(not T.have_reference(), reason="oracle/_ref/genmap_ref not built")


@pytest.mark.parametrize("seed,nchr,length,nfiles", [(5, 3, 1000, 1), (6, 1, 70000, 1), (7, 4, 40000, 1), (8, 2, 3000, 3)])
def test_writer_is_byte_identical_to_reference_index(tmp_path, seed, nchr, length, nfiles):
    files = []
    for f in range(nfiles):
        seqs = T.repeat_rich(seed + 100 * f, nchr, length)
        files.append(("g%02d.fa" % f if nfiles > 1 else "genome.fa", [("f%ds%d" % (f, i), s) for i, s in enumerate(seqs)]))
    src = tmp_path / "fasta"
    src.mkdir()
    for fn, recs in files:
        T.write_fasta(str(src / fn), [c for _, c in recs], names=[n for n, _ in recs])
    ref_dir = str(tmp_path / "ref_index")
    flag = ["-FD", str(src)] if nfiles > 1 else ["-F", str(src / "genome.fa")]
    subprocess.run([T.REF_BIN, "index"] + flag + ["-I", ref_dir], check=True, stdout=subprocess.DEVNULL)
    seqs = [c for _, recs in files for _, c in recs]
    hs = T.HostSim(seqs, with_sa=True)  # the product's own host builder supplies BWT and SA
    ours = T.write_seqan_index(str(tmp_path / "our_index"), files, hs.bwt(False), hs.bwt(True), hs.sa())
    names = sorted(os.listdir(ref_dir))
    assert sorted(os.listdir(ours)) == names
    match, mismatch, errors = filecmp.cmpfiles(ref_dir, ours, names, shallow=False)
    assert not mismatch and not errors, mismatch
    # and the reference maps on it
    out = tmp_path / "out"
    out.mkdir()
    subprocess.run([T.REF_BIN, "map", "-I", ours, "-O", str(out), "-K", "20", "-E", "1", "-r", "-fl"], check=True,
                   stdout=subprocess.DEVNULL)
    stf = np.array([fi for fi, (_, recs) in enumerate(files) for _ in recs], dtype=np.uint32)
    orc = T.Oracle(seqs, seq_to_file=stf)
    for fi, (fn, _) in enumerate(files):
        got = np.fromfile(str(out / (fn[:-3] + ".genmap.freq16")), dtype=np.uint16)
        assert np.array_equal(got, orc.map(20, 1, file_no=fi))


@pytest.mark.parametrize("with_n", [False, True], ids=["dna4", "dna5"])
@pytest.mark.parametrize("seed,nchr,length", [(15, 3, 1000), (16, 2, 70000)])
def test_importing_a_reference_built_index_gives_our_own_blob(tmp_path, seed, nchr, length, with_n):
    """`genmap_ref index` -> gmb_index_import_reference must equal the blob our own builder makes from the FASTA."""
    import genmap_b200
    seqs = T.repeat_rich(seed, nchr, length, with_n=with_n)
    fa = str(tmp_path / "genome.fa")
    T.write_fasta(fa, seqs)
    ref_dir = str(tmp_path / "ref_index")
    subprocess.run([T.REF_BIN, "index", "-F", fa, "-I", ref_dir], check=True, stdout=subprocess.DEVNULL)
    imported = genmap_b200.Index.import_reference_blob(ref_dir)
    ours = genmap_b200.Index.build_blob(seqs, with_sa=False)
    assert imported.tobytes() == ours.tobytes()
    with pytest.raises(genmap_b200.GenmapError):
        genmap_b200.Index.import_reference_blob(str(tmp_path / "nowhere"))


@pytest.mark.parametrize("nfiles,sampling", [(1, 10), (3, 7)])
def test_product_exporter_is_byte_identical_to_reference_index(tmp_path, nfiles, sampling):
    """`genmap index --reference-format` / gmb_blob_export_reference (the product's writer, seqan_export.cpp) against
    the reference's own `genmap index` on the same FASTA input: every fibre identical, and the reference maps on it."""
    import genmap_b200
    from genmap_b200 import _build
    cli = _build.build_cli()
    src = tmp_path / "fasta"
    src.mkdir()
    seqs_all = []
    for f in range(nfiles):
        seqs = T.repeat_rich(21 + f, 2 + f, 30000 + 777 * f)
        seqs_all += seqs
        T.write_fasta(str(src / ("g%02d.fa" % f if nfiles > 1 else "genome.fa")), seqs, names=["f%ds%d extra words" % (f, i) for i in range(len(seqs))])
    flag = ["-FD", str(src)] if nfiles > 1 else ["-F", str(src / "genome.fa")]
    ref_dir, our_dir = str(tmp_path / "ref_index"), str(tmp_path / "our_index")
    subprocess.run([T.REF_BIN, "index"] + flag + ["-I", ref_dir, "-S", str(sampling)], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([cli, "index"] + flag + ["-I", our_dir, "-xh", "-xf", "-S", str(sampling)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    names = sorted(os.listdir(ref_dir))
    assert set(names) <= set(os.listdir(our_dir))
    match, mismatch, errors = filecmp.cmpfiles(ref_dir, our_dir, names, shallow=False)
    assert not mismatch and not errors, (mismatch, errors)
    out = tmp_path / "out"
    out.mkdir()
    subprocess.run([T.REF_BIN, "map", "-I", our_dir, "-O", str(out), "-K", "24", "-E", "1", "-r", "-fl"], check=True, stdout=subprocess.DEVNULL)
    stf = np.array([f for f in range(nfiles) for _ in range(2 + f)], dtype=np.uint32)
    orc = T.Oracle(seqs_all, seq_to_file=stf)
    for f in range(nfiles):
        got = np.fromfile(str(out / (("g%02d" % f if nfiles > 1 else "genome") + ".genmap.freq16")), dtype=np.uint16)
        assert np.array_equal(got, orc.map(24, 1, file_no=f))
    # what the exporter refuses
    with pytest.raises(genmap_b200.GenmapError):
        genmap_b200.Index.export_reference_blob(genmap_b200.Index.build_blob(seqs_all, with_sa=False), str(tmp_path / "x"), ["a;1;b"] * len(seqs_all))


@pytest.mark.parametrize("name,tail,with_n", [("32_16_64", 40, False), ("64_64_64", 70000, False), ("64_64_64_dna5", 70000, True)])
def test_importing_the_reference_other_index_width_classes(tmp_path, name, tail, with_n):
    """More than 65535 sequences put a reference index into its (32,16,64) or (64,64,64) class (src/indexing.hpp:151-170:
    uint32 block counters, 64-bit superblocks, wider suffix-array pairs): they import into the same blob our own builder
    makes, as long as the text has fewer than 2^32 - 1 rows."""
    import genmap_b200
    rng = np.random.default_rng(3)
    seqs = [rng.integers(0, 4, n, dtype=np.uint8) for n in [40] * 69999 + [tail]]
    if with_n:
        seqs[5][7] = 4
        seqs[-1][100:140] = 4
    fa = str(tmp_path / "g.fa")
    T.write_fasta(fa, seqs, names=["s%d" % i for i in range(len(seqs))])
    ref_dir = str(tmp_path / "idx")
    subprocess.run([T.REF_BIN, "index", "-F", fa, "-I", ref_dir], check=True, stdout=subprocess.DEVNULL)
    info = open(os.path.join(ref_dir, "index.info.concat")).read()
    assert "bwt_dimensions:64" in info and ("sa_dimensions_i1:32" in info or "sa_dimensions_i1:64" in info)
    assert ("alphabet_size:5" in info) == with_n
    imported = genmap_b200.Index.import_reference_blob(ref_dir)
    assert imported.tobytes() == genmap_b200.Index.build_blob(seqs, with_sa=False).tobytes()
