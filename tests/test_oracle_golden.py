"""Pins the CPU oracle (oracle/gm_oracle.c) to the reference's own golden vectors
(/root/reference/tests/test_cases, copied to tests/golden/reference_cases) — CPU only."""
import os

import numpy as np
import pytest

import gmtest as T


@pytest.mark.parametrize("case", sorted(T.CASES))
@pytest.mark.parametrize("bits", [16, 8])
def test_oracle_matches_reference_golden(case, bits):
    cfg = T.CASES[case]
    files, sel, folder = T.load_case(case)
    seqs, stf, _ = T.case_layout(files)
    orc = T.Oracle(seqs, seq_to_file=stf)
    ext = "freq16" if bits == 16 else "freq8"
    checked = 0
    for fi, (base, recs) in enumerate(files):
        iv = T.file_intervals(sel, recs)
        gold_path = os.path.join(folder, "raw_" + ext, base + ".genmap." + ext)
        if iv is None:  # no interval in this file -> the reference writes no output
            assert not os.path.exists(gold_path)
            continue
        gold = np.fromfile(gold_path, dtype=np.uint16 if bits == 16 else np.uint8)
        kw = dict(revcompl=cfg["rc"], exclude_pseudo=cfg["ep"], value_bits=bits, file_no=fi, intervals=iv)
        got = orc.map(cfg["K"], cfg["E"], **kw)
        assert np.array_equal(got, gold), (case, base, "fm restatement")
        for infix in range(2 if cfg["E"] == 0 else cfg["E"] + 1 if cfg["E"] == 1 else cfg["E"] + 2, cfg["K"] + 1):
            # overlap-invariance (tests/tests.sh:47-60 re-runs with -xo 1 / -xo 2)
            assert np.array_equal(orc.map(cfg["K"], cfg["E"], infix_len=infix, **kw), gold), (case, base, infix)
        if not cfg["dir"]:
            assert np.array_equal(orc.map(cfg["K"], cfg["E"], copy_shortcut=True, **kw), gold)
        b = T.brute(seqs, cfg["K"], cfg["E"], seq_to_file=stf, **{k: v for k, v in kw.items()})
        assert np.array_equal(b, gold), (case, base, "brute force")
        checked += 1
    assert checked > 0
