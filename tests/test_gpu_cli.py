"""Replays the reference's end-to-end golden matrix (tests/CMakeLists.txt:27-73 x tests/tests.sh) through
the new `genmap` binary on the GPU: index -> map -> diff -r with the expected folder."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

import gmtest as T

pytestmark = pytest.mark.gpu

CASES = sorted(T.CASES)  # all 18 of tests/CMakeLists.txt:56-73 (1c-1g are genomes with N: Dna5 indices)


@pytest.fixture(scope="module")
def genmap():
    from genmap_b200 import _build
    _build.build()
    return _build.build_cli()


VALUE_TYPES = {"map": [], "freq16": ["-fl"], "freq8": ["-fs"]}
FORMATS = {"raw": "-r", "txt": "-t", "wig": "-w", "bed": "-bg"}  # golden folder prefix -> flag


def _replay(genmap, case, tmp_path, host_builder=False, extra=()):
    """tests/tests.sh for one case; one `map` call per value type writes every format at once."""
    cfg = T.CASES[case]
    folder = os.path.join(T.GOLDEN, "reference_cases", "case_" + case)
    idx = str(tmp_path / ("index" + "".join(extra) + ("h" if host_builder else "")))
    src = ["-FD", folder] if cfg["dir"] else ["-F", os.path.join(folder, "genome.fa")]
    r = subprocess.run([genmap, "index"] + src + ["-I", idx] + (["-xh"] if host_builder else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    base_flags = ["-K", str(cfg["K"]), "-E", str(cfg["E"])] + ([] if cfg["rc"] else ["-nc"]) + (["-ep"] if cfg["ep"] else [])
    if os.path.exists(os.path.join(folder, "subset.bed")):
        base_flags += ["-S", os.path.join(folder, "subset.bed")]
    n = 0
    for vt, vflags in VALUE_TYPES.items():
        golden = {fmt: os.path.join(folder, "%s_%s" % (fmt, vt)) for fmt in FORMATS}
        golden = {fmt: d for fmt, d in golden.items() if os.path.isdir(d)}
        if not golden:
            continue
        out = tmp_path / ("out_%s_%s%s" % (vt, "".join(extra), "h" if host_builder else ""))
        out.mkdir()
        r = subprocess.run([genmap, "map", "-I", idx, "-O", str(out)] + base_flags + vflags + [FORMATS[f] for f in golden] + list(extra),
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        expected = set()
        for fmt, d in golden.items():
            names = os.listdir(d)
            expected.update(names)
            match, mismatch, errors = filecmp.cmpfiles(d, str(out), names, shallow=False)
            assert not mismatch and not errors, (case, vt, fmt, mismatch, errors)
            n += len(match)
        assert set(os.listdir(str(out))) == expected, (case, vt, set(os.listdir(str(out))) ^ expected)
    # only track formats: the runs come from the GPU (gmb_map_runs), the vector never reaches the host
    for vt, vflags in VALUE_TYPES.items():
        golden = {fmt: os.path.join(folder, "%s_%s" % (fmt, vt)) for fmt in ("wig", "bed")}
        golden = {fmt: d for fmt, d in golden.items() if os.path.isdir(d)}
        if not golden:
            continue
        out = tmp_path / ("out_runs_%s_%s%s" % (vt, "".join(extra), "h" if host_builder else ""))
        out.mkdir()
        r = subprocess.run([genmap, "map", "-I", idx, "-O", str(out)] + base_flags + vflags + [FORMATS[f] for f in golden] + list(extra),
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        for fmt, d in golden.items():
            names = os.listdir(d)
            match, mismatch, errors = filecmp.cmpfiles(d, str(out), names, shallow=False)
            assert not mismatch and not errors, (case, vt, fmt, "device runs", mismatch, errors)
    # csv (`-d`, tests/CMakeLists.txt:52-53): its own map call, like the reference's test matrix
    golden = os.path.join(folder, "csv")
    out = tmp_path / ("out_csv_%s%s" % ("".join(extra), "h" if host_builder else ""))
    out.mkdir()
    r = subprocess.run([genmap, "map", "-I", idx, "-O", str(out)] + base_flags + ["-d"] + list(extra), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    names = os.listdir(golden)
    match, mismatch, errors = filecmp.cmpfiles(golden, str(out), names, shallow=False)
    assert not mismatch and not errors, (case, "csv", mismatch, errors)
    assert set(os.listdir(str(out))) == set(names)
    assert n > 0


@pytest.mark.parametrize("case", CASES)
def test_cli_reproduces_reference_golden_directories(genmap, case, tmp_path):
    _replay(genmap, case, tmp_path)


@pytest.mark.parametrize("case", ["1d", "1g", "2b", "3d"])
def test_cli_golden_with_host_built_index_and_overlap_flag(genmap, case, tmp_path):
    _replay(genmap, case, tmp_path, host_builder=True)
    cfg = T.CASES[case]
    if min(cfg["K"] - 1, cfg["K"] - cfg["E"] - 2) >= 1:  # tests.sh:46-47: 1e/1f/1g (K=3, E=1) allow no larger overlap
        _replay(genmap, case, tmp_path, extra=("-xo", "1"))  # tests.sh:47-60 re-runs with -xo: results must not change


def test_cli_output_prefix_and_verbose(genmap, tmp_path):
    folder = os.path.join(T.GOLDEN, "reference_cases", "case_2b")
    idx = str(tmp_path / "index")
    assert subprocess.run([genmap, "index", "-F", os.path.join(folder, "genome.fa"), "-I", idx]).returncode == 0
    r = subprocess.run([genmap, "map", "-I", idx, "-O", str(tmp_path / "myprefix"), "-K", "4", "-E", "0", "-r", "-fl", "-t", "-v"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Mappability computed in" in r.stdout
    assert filecmp.cmp(str(tmp_path / "myprefix.freq16"), os.path.join(folder, "raw_freq16", "genome.genmap.freq16"), shallow=False)
    assert os.path.exists(str(tmp_path / "myprefix.txt"))


def test_cli_maps_on_an_index_written_by_the_reference(genmap, tmp_path):
    """Pre-built GenMap indices are usable as they are: `genmap_ref index` -> our `genmap map`."""
    if not T.have_reference():
        pytest.skip("oracle/_ref/genmap_ref not present")
    for case in ("2b", "3b", "1f"):  # 1f: a genome with N, i.e. a Dna5 index of the reference
        cfg = T.CASES[case]
        folder = os.path.join(T.GOLDEN, "reference_cases", "case_" + case)
        idx = str(tmp_path / ("refindex_" + case))
        src = ["-FD", folder] if cfg["dir"] else ["-F", os.path.join(folder, "genome.fa")]
        subprocess.run([T.REF_BIN, "index"] + src + ["-I", idx], check=True, stdout=subprocess.DEVNULL)
        out = tmp_path / ("out_" + case)
        out.mkdir()
        r = subprocess.run([genmap, "map", "-I", idx, "-O", str(out), "-K", str(cfg["K"]), "-E", str(cfg["E"]), "-r", "-fl", "-w"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        for sub in ("raw_freq16", "wig_freq16"):
            names = os.listdir(os.path.join(folder, sub))
            match, mismatch, errors = filecmp.cmpfiles(os.path.join(folder, sub), str(out), names, shallow=False)
            assert not mismatch and not errors, (case, sub, mismatch, errors)


@pytest.mark.parametrize("with_n", [False, True], ids=["dna4", "dna5"])
def test_cli_multi_gpu_sharding_gives_identical_files(genmap, tmp_path, with_n):
    from genmap_b200 import _lib
    if _lib.lib().gmb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import genmap_b200 as gm
    fa = str(tmp_path / "g.fa")
    seqs = gm.synth_genome(600_000, 3, 3)
    if with_n:  # assembly gaps, one of them across the middle of the text where the two GPUs' slices meet: every GPU
        seqs[0][50_000:60_000] = 4  # builds the N pass for the whole index and applies it to its own positions
        seqs[1][95_000:105_000] = 4
        seqs[2][1000:1003] = 4
    T.write_fasta(fa, seqs)
    idx = str(tmp_path / "index")
    assert subprocess.run([genmap, "index", "-F", fa, "-I", idx]).returncode == 0
    outs = []
    for n in (1, 2):
        out = tmp_path / ("out%d" % n)
        out.mkdir()
        r = subprocess.run([genmap, "map", "-I", idx, "-O", str(out), "-K", "30", "-E", "1", "-r", "-fl", "-bg", "-xg", str(n)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(out)
        # device runs, one slice per GPU merged on the host: the same bedgraph
        out2 = tmp_path / ("out%d_runs" % n)
        out2.mkdir()
        r = subprocess.run([genmap, "map", "-I", idx, "-O", str(out2), "-K", "30", "-E", "1", "-fl", "-bg", "-w", "-xg", str(n)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert filecmp.cmp(str(out / "g.genmap.bedgraph"), str(out2 / "g.genmap.bedgraph"), shallow=False)
    for name in os.listdir(str(outs[0])):
        assert filecmp.cmp(str(outs[0] / name), str(outs[1] / name), shallow=False), name
    assert np.array_equal(np.fromfile(str(outs[1] / "g.genmap.freq16"), dtype=np.uint16), T.Oracle(seqs).map(30, 1))


def test_cli_csv_matches_the_reference_binary(genmap, tmp_path):
    """`-d` next to other formats, E > 0, multi-sequence, with and without a selection: byte-identical csv."""
    if not T.have_reference():
        pytest.skip("oracle/_ref/genmap_ref not present")
    import genmap_b200 as gm
    fa = str(tmp_path / "g.fa")
    T.write_fasta(fa, T.repeat_rich(21, 3, 4000))
    bed = str(tmp_path / "sel.bed")
    with open(bed, "w") as f:
        f.write("chr1\t100\t900\nchr3\t3500\t4000\n")
    ref_idx, idx = str(tmp_path / "ref_index"), str(tmp_path / "index")
    subprocess.run([T.REF_BIN, "index", "-F", fa, "-I", ref_idx], check=True, stdout=subprocess.DEVNULL)
    assert subprocess.run([genmap, "index", "-F", fa, "-I", idx]).returncode == 0
    for tag, flags in (("e1", ["-K", "20", "-E", "1"]), ("e2nc", ["-K", "24", "-E", "2", "-nc"]), ("sel", ["-K", "16", "-E", "1", "-S", bed])):
        outs = []
        for binary, index in ((T.REF_BIN, ref_idx), (genmap, idx)):
            out = tmp_path / ("out_%s_%d" % (tag, len(outs)))
            out.mkdir()
            r = subprocess.run([binary, "map", "-I", index, "-O", str(out)] + flags + ["-d", "-r", "-fl"], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            outs.append(out)
        for name in ("g.genmap.csv", "g.genmap.freq16"):
            assert filecmp.cmp(str(outs[0] / name), str(outs[1] / name), shallow=False), (tag, name)
        assert os.path.getsize(str(outs[1] / "g.genmap.csv")) > 10000
