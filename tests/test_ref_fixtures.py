"""Outputs of the unmodified reference binary on seeded genomes (tests/golden/ref_synth.npz, made by
tests/golden/make_fixtures.py) pin the oracle and the host-compiled kernel state machine — CPU only."""
import os

import numpy as np
import pytest

import gmtest as T

FIX = np.load(os.path.join(T.GOLDEN, "ref_synth.npz"))
CASES = [str(c) for c in FIX["cases"]]


def load_fixture(line):
    name, K, E, flags, bits, nfiles = line.split("|")
    K, E, bits, nfiles = int(K), int(E), int(bits), int(nfiles)
    files = []
    for fi in range(nfiles):
        seqs, si = [], 0
        while "%s/in/%d/%d" % (name, fi, si) in FIX:
            seqs.append(FIX["%s/in/%d/%d" % (name, fi, si)]); si += 1
        files.append(seqs)
    outs = [FIX["%s/out/%d" % (name, fi)] for fi in range(nfiles)]
    seqs = [s for f in files for s in f]
    stf = np.array([fi for fi, f in enumerate(files) for _ in f], dtype=np.uint32)
    return name, K, E, flags.split(), bits, seqs, stf, outs


@pytest.mark.parametrize("line", CASES, ids=[c.split("|")[0] for c in CASES])
def test_oracle_matches_reference_binary(line):
    name, K, E, flags, bits, seqs, stf, outs = load_fixture(line)
    orc = T.Oracle(seqs, seq_to_file=stf)
    for fi, gold in enumerate(outs):
        got = orc.map(K, E, revcompl="-nc" not in flags, exclude_pseudo="-ep" in flags, value_bits=bits, file_no=fi)
        assert np.array_equal(got, gold), (name, fi)


@pytest.mark.parametrize("line", [c for c in CASES if c.startswith(("dna4_K12", "dna5_K20_E1", "short", "multi_K25_E1"))],
                         ids=lambda c: c.split("|")[0])
def test_brute_force_matches_reference_binary(line):
    name, K, E, flags, bits, seqs, stf, outs = load_fixture(line)
    for fi, gold in enumerate(outs):
        got = T.brute(seqs, K, E, revcompl="-nc" not in flags, exclude_pseudo="-ep" in flags, value_bits=bits,
                      seq_to_file=stf, file_no=fi)
        assert np.array_equal(got, gold), (name, fi)


@pytest.mark.parametrize("line", [c for c in CASES if not c.startswith("dna5") and "-ep" not in c],
                         ids=lambda c: c.split("|")[0])
def test_kernel_state_machine_matches_reference_binary(line):
    name, K, E, flags, bits, seqs, stf, outs = load_fixture(line)
    hs = T.HostSim(seqs)
    for fi, gold in enumerate(outs):
        got = hs.map(K, E, revcompl="-nc" not in flags, value_bits=bits, seq_to_file=stf, file_no=fi)
        assert np.array_equal(got, gold), (name, fi)
