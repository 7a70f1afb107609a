"""CPU-side checks of the product library: it loads, exports every symbol include/genmap_b200.h
declares, its host index builder is deterministic, and the compute entry points refuse to run
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import gmtest as T

import genmap_b200
from genmap_b200 import _build, _lib


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(T.ROOT, "include", "genmap_b200.h")).read()
    declared = set(re.findall(r"\b(gmb_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.gmb_version()


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.GmbParams) == 32
    assert ctypes.sizeof(_lib.GmbIndexInfo) == 64
    assert ctypes.sizeof(_lib.GmbMapStats) == 144


def test_host_builder_blob_is_deterministic_and_matches_hostsim(lib):
    seqs = T.repeat_rich(5, 3, 2500)
    a = genmap_b200.Index.build_blob(seqs, with_sa=True)
    b = genmap_b200.Index.build_blob(seqs, with_sa=True)
    assert a.tobytes() == b.tobytes()
    hs = T.HostSim(seqs, with_sa=True)
    assert a.tobytes() == hs.blob_bytes()


def test_builder_rejects_bad_input(lib):
    with pytest.raises(genmap_b200.GenmapError) as e:
        genmap_b200.Index.build_blob([np.array([0, 1, 5, 2], np.uint8)])
    assert e.value.code == _lib.GMB_ERR_ARG and "invalid base code" in str(e.value)
    with pytest.raises(genmap_b200.GenmapError):
        genmap_b200.Index.build_blob([np.array([0, 1], np.uint8), np.zeros(0, np.uint8)])


def test_n_makes_a_dna5_blob(lib):
    # src/indexing.hpp:459-473: one N anywhere switches the whole index to the Dna5 alphabet
    hdr4 = np.frombuffer(genmap_b200.Index.build_blob([np.array([0, 1, 3, 2] * 8, np.uint8)]), np.uint32, 4)
    hdr5 = np.frombuffer(genmap_b200.Index.build_blob([np.array([0, 1, 4, 2] * 8, np.uint8)]), np.uint32, 4)
    assert hdr4[3] == 4 and hdr5[3] == 5


def test_no_cpu_fallback(lib):
    if lib.gmb_device_count() > 0:
        pytest.skip("a GPU is present")
    blob = genmap_b200.Index.build_blob(T.repeat_rich(5, 1, 500))
    with pytest.raises(genmap_b200.GenmapError) as e:
        genmap_b200.Index.from_blob(blob)
    assert e.value.code == _lib.GMB_ERR_CUDA
    with pytest.raises(genmap_b200.GenmapError) as e:
        genmap_b200.Index.build(T.repeat_rich(5, 1, 500), on_gpu=True)
    assert e.value.code == _lib.GMB_ERR_CUDA


def test_bad_blob_is_rejected(lib):
    h = ctypes.c_void_p()
    junk = np.zeros(4096, np.uint8)
    rc = lib.gmb_index_from_blob(ctypes.c_void_p(junk.ctypes.data), junk.nbytes, 0, ctypes.byref(h))
    assert rc == _lib.GMB_ERR_IO and b"magic" in lib.gmb_last_error()


def test_synth_generator_is_the_frozen_one():
    a = genmap_b200.synth_genome(100000, 2, 42)
    b = T.synth(100000, 2, 42)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_corrupt_or_truncated_blobs_are_rejected_before_they_reach_a_kernel(lib):
    """section offsets, total size and cumulative counts are checked against the layout of an index of
    the header's own sizes (validate_header), not just for being inside the blob."""
    blob = genmap_b200.Index.build_blob(T.repeat_rich(5, 3, 2500), with_sa=True)
    words = blob.view(np.uint64)

    def rejected(b):
        with pytest.raises(genmap_b200.GenmapError) as e:
            genmap_b200.Index.from_blob(b)
        return e.value.code == _lib.GMB_ERR_IO

    assert rejected(blob[:len(blob) - 256].copy())                     # truncated
    hdr_words = 256 // 8
    hit = 0
    for i in range(hdr_words):                                          # any changed size, count or offset
        if words[i] == 0:
            continue
        bad = blob.copy()
        bad.view(np.uint64)[i] += 256
        with pytest.raises(genmap_b200.GenmapError) as e:
            genmap_b200.Index.from_blob(bad)
        hit += e.value.code == _lib.GMB_ERR_IO
    assert hit >= 15
    if lib.gmb_device_count() == 0:                                     # the intact blob gets as far as "no device" here
        with pytest.raises(genmap_b200.GenmapError) as e:
            genmap_b200.Index.from_blob(blob)
        assert e.value.code == _lib.GMB_ERR_CUDA
