"""CPU checks of the `genmap` command line: the writers reproduce every golden output flavour of the
reference's test cases byte for byte (rendered from the golden raw vectors, so no GPU is needed),
`genmap index` works with the host builder, user errors give exit code 1 with the reference's messages."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

import gmtest as T
from genmap_b200 import _build

FLAVOURS = {"raw_map": ["-r"], "raw_freq8": ["-r", "-fs"], "raw_freq16": ["-r", "-fl"], "txt_map": ["-t"],
            "txt_freq16": ["-t", "-fl"], "txt_freq8": ["-t", "-fs"], "wig_map": ["-w"], "wig_freq16": ["-w", "-fl"],
            "bed_map": ["-bg"], "bed_freq16": ["-bg", "-fl"]}


@pytest.fixture(scope="module")
def genmap():
    _build.build()
    return _build.build_cli()


def run(genmap, *args):
    return subprocess.run([genmap] + [str(a) for a in args], capture_output=True, text=True)


def make_index(genmap, case, tmp, host=True):
    files, sel, folder = T.load_case(case)
    idx = os.path.join(tmp, "index")
    if T.CASES[case]["dir"]:
        r = run(genmap, "index", "-FD", folder, "-I", idx, *(["-xh"] if host else []))
    else:
        r = run(genmap, "index", "-F", os.path.join(folder, "genome.fa"), "-I", idx, *(["-xh"] if host else []))
    return r, idx, files, sel, folder


@pytest.mark.parametrize("case", sorted(T.CASES))
def test_index_and_writers_reproduce_golden_outputs(genmap, case, tmp_path):
    r, idx, files, sel, folder = make_index(genmap, case, str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "Index created successfully." in r.stdout
    ids = open(os.path.join(idx, "index.ids")).read().split("\n")[:-1]
    want = ["%s.fa;%d;%s" % (base, len(c), name) for base, recs in files for name, c in recs]
    assert ids == want
    n_checked = 0
    # every flavour through the vector writers; the track formats once more from a run list shaped like the
    # one the GPU returns (gmb_map_runs), which is what `map -w/-bg/-b` without -r/-t uses
    todo = [(flav, flags, "") for flav, flags in FLAVOURS.items()]
    todo += [(flav, flags + ["-xr"], "_runs") for flav, flags in FLAVOURS.items() if flav.startswith(("wig", "bed"))]
    for flav, flags, tag in todo:
        gold_dir = os.path.join(folder, flav)
        if not os.path.isdir(gold_dir):
            continue
        out = tmp_path / (flav + tag)
        out.mkdir()
        for fi, (base, recs) in enumerate(files):
            # render from the golden raw vector of the same value type (freq16 for the float outputs)
            src = os.path.join(folder, "raw_freq8" if "-fs" in flags else "raw_freq16",
                               base + ".genmap." + ("freq8" if "-fs" in flags else "freq16"))
            if not os.path.exists(src):
                continue
            rr = run(genmap, "render", "-I", os.path.join(idx, "index.ids"), "-C", src, "-N", fi,
                     "-O", str(out / (base + ".genmap")), *flags)
            assert rr.returncode == 0, rr.stderr
        cmp = filecmp.dircmp(gold_dir, str(out))
        assert not cmp.left_only and not cmp.right_only and not cmp.funny_files, (flav, cmp.left_only, cmp.right_only)
        match, mismatch, errors = filecmp.cmpfiles(gold_dir, str(out), cmp.common_files, shallow=False)
        assert not mismatch and not errors, (case, flav, mismatch)
        n_checked += len(match)
    assert n_checked > 0


def test_index_dna5_and_existing_directory(genmap, tmp_path):
    (tmp_path / "n").mkdir()
    r, idx, *_ = make_index(genmap, "1c", str(tmp_path / "n"))
    assert r.returncode == 0 and "alphabet_size:5" in open(os.path.join(idx, "index.info")).read()
    r, idx, *_ = make_index(genmap, "1a", str(tmp_path))
    assert r.returncode == 0 and "alphabet_size:4" in open(os.path.join(idx, "index.info")).read()
    r, idx, *_ = make_index(genmap, "1a", str(tmp_path))
    assert r.returncode == 1 and "already exists" in r.stderr


def test_cli_user_errors_exit_1(genmap, tmp_path):
    r, idx, *_ = make_index(genmap, "2a", str(tmp_path))
    out = tmp_path / "out"
    out.mkdir()
    cases = [
        (["map", "-I", idx, "-O", out, "-r"], "Missing value for option: -K, --length"),
        (["map", "-I", idx, "-O", out, "-K", 30, "-E", 5, "-r"], "E > 4 not yet supported."),
        (["map", "-I", idx, "-O", out, "-K", 30], "Please choose at least one output format"),
        (["map", "-I", idx, "-O", out, "-K", 30, "-r", "-fs", "-fl"], "Cannot use both --frequency-small and --frequency-large"),
        (["map", "-I", idx, "-O", tmp_path / "nodir" / "x", "-K", 30, "-r"], "does not exist"),
        (["map", "-I", tmp_path / "noindex", "-O", out, "-K", 30, "-r"], "index"),
        (["map", "-I", idx, "-O", out, "-K", 3, "-E", 2, "-r"], "K must be at least E + 2"),
        (["map", "-I", idx, "-O", out, "-K", 30, "-E", 1, "-xo", 29, "-r"], "overlap cannot be larger"),
        (["index", "-I", tmp_path / "i2"], "You forgot to specify --fasta-file or --fasta-directory"),
        (["frobnicate"], "not in the list of allowed values"),
    ]
    for args, msg in cases:
        r = run(genmap, *args)
        assert r.returncode == 1, (args, r.stdout, r.stderr)
        assert msg in r.stderr + r.stdout, (args, r.stderr)
    assert run(genmap, "--help").returncode == 0 and run(genmap, "--version").returncode == 0
    assert run(genmap, "map", "--help").returncode == 0


def test_map_needs_a_gpu(genmap, tmp_path):
    from genmap_b200 import _lib
    if _lib.lib().gmb_device_count() > 0:
        pytest.skip("a GPU is present")
    r, idx, *_ = make_index(genmap, "2a", str(tmp_path))
    out = tmp_path / "out"
    out.mkdir()
    r = run(genmap, "map", "-I", idx, "-O", out, "-K", 4, "-r", "-fl")
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_txt_writer_chunks_and_threads(genmap, tmp_path):
    """Sequences longer than one formatting chunk (2 Mi values), formatted by several threads: same text as a
    straight rendering (src/output.hpp:40-69), for frequencies and for their inverses."""
    rng = np.random.default_rng(5)
    lens = [5_000_003, 2_097_152, 7]
    c = rng.integers(0, 40, sum(lens)).astype(np.uint16)
    c[rng.integers(0, len(c), 1000)] = 65535
    c.tofile(str(tmp_path / "c.freq16"))
    with open(str(tmp_path / "index.ids"), "w") as f:
        for i, n in enumerate(lens):
            f.write("g.fa;%d;s%d\n" % (n, i))
    for flags, fmt in ((["-fl"], lambda v: "%d" % v), ([], lambda v: "%g" % np.float32(1.0 / np.float32(v)) if v else "0")):
        r = run(genmap, "render", "-I", tmp_path / "index.ids", "-C", tmp_path / "c.freq16", "-N", 0, "-O", tmp_path / "o", "-t", *flags)
        assert r.returncode == 0, r.stderr
        lut = {int(v): fmt(int(v)) for v in np.unique(c)}
        want, b = [], 0
        for i, n in enumerate(lens):
            want.append(">s%d\n%s\n" % (i, " ".join(lut[int(v)] for v in c[b:b + n])))
            b += n
        assert open(str(tmp_path / "o.txt")).read() == "".join(want)


def test_track_writers_from_run_lists_in_chunks_and_threads(genmap, tmp_path):
    """More runs than one formatting chunk (2^18) per sequence, zero runs at chunk borders, equal run lengths across
    borders (the wig header rule): the threaded run-list writers give the files of the serial vector writers."""
    rng = np.random.default_rng(6)
    lens = [1_300_000, 700_001, 5]
    c = rng.integers(0, 3, sum(lens)).astype(np.uint16)            # short runs, many zeros
    c[200_000:700_000] = np.repeat(rng.integers(0, 4, 250_000), 2)  # long stretch of runs of length 2
    c[rng.integers(0, len(c), 500)] = 65535
    c.tofile(str(tmp_path / "c.freq16"))
    with open(str(tmp_path / "index.ids"), "w") as f:
        for i, n in enumerate(lens):
            f.write("g.fa;%d;s%d\n" % (n, i))
    for flags in (["-fl"], []):
        outs = []
        for tag, extra in (("vec", []), ("runs", ["-xr"]), ("serial", ["-T", "1"]), ("many", ["-T", "7", "-r"])):
            out = tmp_path / ("o_%s_%d" % (tag, len(flags)))
            r = run(genmap, "render", "-I", tmp_path / "index.ids", "-C", tmp_path / "c.freq16", "-N", 0, "-O", out, "-w", "-bg", "-b",
                    *flags, *extra)
            assert r.returncode == 0, r.stderr
            outs.append(str(out))
        for ext in (".wig", ".bedgraph", ".bed", ".chrom.sizes"):
            for other in outs[1:]:  # threaded scan of the vector, run list, the serial scan, 7 threads: the same files
                assert filecmp.cmp(outs[0] + ext, other + ext, shallow=False), (flags, ext, other)
        raw = np.fromfile(outs[3] + (".freq16" if flags else ".map"), dtype=np.uint16 if flags else np.float32)  # threaded raw writer
        want = c if flags else np.where(c != 0, np.float32(1.0) / np.maximum(c, 1).astype(np.float32), np.float32(0)).astype(np.float32)
        assert np.array_equal(raw, want)
        assert os.path.getsize(outs[0] + ".wig") > 5_000_000


def test_index_reads_fastq_and_awkward_fasta(genmap, tmp_path):
    """-FD accepts .fastq like the reference's SeqFileIn; CRLF line ends, blank lines, lower case and
    spaces inside FASTA sequence lines give the same index as the clean file."""
    seqs = T.repeat_rich(4, 3, 700)
    txt = ["".join("ACGT"[c] for c in s) for s in seqs]
    d = tmp_path / "fq"
    d.mkdir()
    with open(d / "reads.fastq", "w") as f:
        for i, t in enumerate(txt):
            f.write("@r%d some comment\n%s\n%s\n+\n%s\n%s\n" % (i, t[:300], t[300:], "I" * 300, "@" * (len(t) - 300)))  # multi-line, '@' in the qualities
    clean, messy = tmp_path / "clean.fa", tmp_path / "messy.fa"
    with open(clean, "w") as f:
        for i, t in enumerate(txt):
            f.write(">r%d some comment\n%s\n" % (i, t))
    with open(messy, "wb") as f:
        f.write(b"\r\n")
        for i, t in enumerate(txt):
            f.write((">r%d some comment\r\n" % i).encode())
            for k in range(0, len(t), 61):
                f.write((t[k:k + 30].lower() + " " + t[k + 30:k + 61] + "\r\n\r\n").encode())
    blobs = {}
    for name, flag, src in (("fastq", "-FD", d), ("clean", "-F", clean), ("messy", "-F", messy)):
        idx = tmp_path / ("idx_" + name)
        r = run(genmap, "index", flag, src, "-I", idx, "-xh")
        assert r.returncode == 0, r.stderr
        blobs[name] = open(idx / "index.gmb", "rb").read()
        ids = open(idx / "index.ids").read().split("\n")
        assert ids[0].split(";")[1:] == ["700", "r0"], ids[0]
    assert blobs["fastq"] == blobs["clean"] == blobs["messy"]
