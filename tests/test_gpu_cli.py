"""Replays the reference's end-to-end golden matrix (tests/CMakeLists.txt:27-73 x tests/tests.sh) through
the new `genmap` binary on the GPU: index -> map -> diff -r with the expected folder."""
import filecmp
import os
import subprocess

import pytest

import gmtest as T
from test_cli_cpu import FLAVOURS

pytestmark = pytest.mark.gpu

# Dna5 genomes (1c-1g) are not on the GPU path yet
CASES = ["1a", "1b", "2a", "2b", "2c", "2d", "2e", "3a", "3b", "3c", "3d", "3e", "3f"]


@pytest.fixture(scope="module")
def genmap():
    from genmap_b200 import _build
    _build.build()
    return _build.build_cli()


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("builder", ["gpu", "host"])
def test_cli_reproduces_reference_golden_directories(genmap, case, builder, tmp_path):
    cfg = T.CASES[case]
    folder = os.path.join(T.GOLDEN, "reference_cases", "case_" + case)
    idx = str(tmp_path / "index")
    src = ["-FD", folder] if cfg["dir"] else ["-F", os.path.join(folder, "genome.fa")]
    r = subprocess.run([genmap, "index"] + src + ["-I", idx] + (["-xh"] if builder == "host" else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    base_flags = ["-K", str(cfg["K"]), "-E", str(cfg["E"])] + ([] if cfg["rc"] else ["-nc"]) + (["-ep"] if cfg["ep"] else [])
    if os.path.exists(os.path.join(folder, "subset.bed")):
        base_flags += ["-S", os.path.join(folder, "subset.bed")]
    n = 0
    for flav, flags in FLAVOURS.items():
        gold = os.path.join(folder, flav)
        if not os.path.isdir(gold):
            continue
        for extra in ([], ["-xo", "1"]):  # tests.sh:47-60 re-runs with -xo: results must not change
            out = tmp_path / (flav + "_" + "".join(extra))
            out.mkdir()
            r = subprocess.run([genmap, "map", "-I", idx, "-O", str(out)] + base_flags + flags + extra, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            cmp = filecmp.dircmp(gold, str(out))
            assert not cmp.left_only and not cmp.right_only, (case, flav, cmp.left_only, cmp.right_only)
            match, mismatch, errors = filecmp.cmpfiles(gold, str(out), cmp.common_files, shallow=False)
            assert not mismatch and not errors, (case, flav, mismatch)
            n += len(match)
    assert n > 0


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
