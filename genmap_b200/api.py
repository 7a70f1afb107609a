"""Host-side mirror of the reference's interface for the map path.

The reference has no library API: `genmap map` opens the index once (src/mappability.hpp:221-223) and
calls, per FASTA file, run(index, text, ...) -> computeMappability<E>(index, text, c, searchParams, ...)
(src/mappability.hpp:157-189, src/algo.hpp:405-410).  `Index.open/build` and `compute_mappability`
mirror those two steps with the same argument meaning (file-local positions, counts over the whole
index, uint8/uint16 value types, -nc / -ep switches, selection intervals) and the same error
behaviour (E > 4 is rejected, src/mappability.hpp:187).  Everything is a thin ctypes call into the
C ABI (include/genmap_b200.h); the arithmetic happens in the CUDA kernels.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import GenmapError, GmbIndexInfo, GmbLocations, GmbMapStats, GmbParams, GmbRuns, check


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _concat(seqs):
    codes = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]))
    limits = np.zeros(len(seqs) + 1, dtype=np.uint64)
    limits[1:] = np.cumsum([len(s) for s in seqs])
    return codes, limits


class SearchParams:
    """src/common.hpp:67-74 (length, revCompl, excludePseudo) + Options.errors / value type."""

    def __init__(self, length, errors=0, rev_compl=True, exclude_pseudo=False, value_bits=16, block_kmers=0):
        self.length, self.errors, self.rev_compl = int(length), int(errors), bool(rev_compl)
        self.exclude_pseudo, self.value_bits = bool(exclude_pseudo), int(value_bits)
        self.block_kmers = int(block_kmers)  # k-mers per block (K - overlap + 1 in the reference); 0 = default


class Index:
    """An FM index resident in the HBM of one GPU."""

    def __init__(self, handle, seq_to_file=None):
        self._h = handle
        info = GmbIndexInfo()
        check(_lib.lib().gmb_index_get_info(self._h, ctypes.byref(info)))
        self.info = info
        self.n_text, self.n_seq, self.device = int(info.n_text), int(info.n_seq), int(info.device)
        self.seq_to_file = None if seq_to_file is None else np.ascontiguousarray(seq_to_file, dtype=np.uint32)
        self.limits = None
        self.build_timings_ms = None

    # ---- construction ---------------------------------------------------------------------------
    @staticmethod
    def build_blob(seqs, with_sa=False, on_gpu=False, device=0):
        """-> index blob as a uint8 array (host SA-IS builder, or the GPU builder copied back)."""
        codes, limits = _concat(seqs)
        blob, nbytes = ctypes.c_void_p(), ctypes.c_uint64()
        flags = (_lib.GMB_BUILD_WITH_SA if with_sa else 0) | (_lib.GMB_BUILD_ON_GPU if on_gpu else 0)
        check(_lib.lib().gmb_index_build(_ptr(codes), _ptr(limits), len(seqs), flags, device,
                                         ctypes.byref(blob), ctypes.byref(nbytes)))
        out = np.frombuffer(ctypes.string_at(blob, nbytes.value), dtype=np.uint8).copy()
        _lib.lib().gmb_blob_free(blob)
        return out

    @staticmethod
    def import_reference_blob(directory):
        """-> index blob converted from an index directory written by the reference's `genmap index`."""
        blob, nbytes = ctypes.c_void_p(), ctypes.c_uint64()
        check(_lib.lib().gmb_index_import_reference(str(directory).encode(), ctypes.byref(blob), ctypes.byref(nbytes)))
        out = np.frombuffer(ctypes.string_at(blob, nbytes.value), dtype=np.uint8).copy()
        _lib.lib().gmb_blob_free(blob)
        return out

    @staticmethod
    def export_reference_blob(blob, directory, ids, fasta_directory=False, sampling=10):
        """Write a blob built with with_sa=True as an index directory in the reference's own format (what its `map`
        opens).  ids: one "file;length;name" string per sequence."""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        import os
        os.makedirs(str(directory), exist_ok=True)
        arr = (ctypes.c_char_p * len(ids))(*[s.encode() for s in ids])
        check(_lib.lib().gmb_blob_export_reference(_ptr(blob), blob.nbytes, str(directory).encode(), arr, len(ids),
                                                   int(fasta_directory), int(sampling)))

    @classmethod
    def build(cls, seqs, device=0, with_sa=False, on_gpu=True, seq_to_file=None):
        """Index the sequences (uint8 codes 0..3) and leave the index in HBM of `device`."""
        codes, limits = _concat(seqs)
        h = ctypes.c_void_p()
        if on_gpu:
            tm = (ctypes.c_double * 4)()
            check(_lib.lib().gmb_index_build_device(_ptr(codes), _ptr(limits), len(seqs),
                                                    _lib.GMB_BUILD_WITH_SA if with_sa else 0, device, ctypes.byref(h), tm))
            ix = cls(h, seq_to_file)
            ix.build_timings_ms = dict(h2d=tm[0], sort=tm[1], pack=tm[2], total=tm[3])
        else:
            blob = cls.build_blob(seqs, with_sa=with_sa)
            check(_lib.lib().gmb_index_from_blob(_ptr(blob), blob.nbytes, device, ctypes.byref(h)))
            ix = cls(h, seq_to_file)
        ix.limits = limits
        return ix

    @classmethod
    def from_blob(cls, blob, device=0, seq_to_file=None):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        h = ctypes.c_void_p()
        check(_lib.lib().gmb_index_from_blob(_ptr(blob), blob.nbytes, device, ctypes.byref(h)))
        return cls(h, seq_to_file)

    @classmethod
    def adopt_device(cls, device_ptr, nbytes, device=0, seq_to_file=None):
        """Wrap a blob that already sits in device memory (e.g. a torch tensor after dist.broadcast)."""
        h = ctypes.c_void_p()
        check(_lib.lib().gmb_index_adopt_device(ctypes.c_void_p(device_ptr), nbytes, device, ctypes.byref(h)))
        return cls(h, seq_to_file)

    @classmethod
    def open(cls, directory, device=0, seq_to_file=None):
        h = ctypes.c_void_p()
        check(_lib.lib().gmb_index_open(str(directory).encode(), device, ctypes.byref(h)))
        return cls(h, seq_to_file)

    def replicate(self, device):
        """A copy of this index in the HBM of another GPU of the node (peer-to-peer copy over NVLink)."""
        h = ctypes.c_void_p()
        check(_lib.lib().gmb_index_replicate(self._h, int(device), ctypes.byref(h)))
        ix = Index(h, self.seq_to_file)
        ix.limits = self.limits
        return ix

    def close(self):
        if self._h:
            _lib.lib().gmb_index_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the hot path ------------------------------------------------------------------------------
    def _file_args(self, text_begin, text_len, chrom_cum, intervals):
        if text_len is None:
            text_begin, text_len = 0, self.n_text
        if chrom_cum is None:
            if self.limits is None or (text_begin, text_len) != (0, self.n_text):
                raise ValueError("chrom_cum_lengths is required")
            chrom_cum = self.limits
        chrom_cum = np.ascontiguousarray(chrom_cum, dtype=np.uint64)
        iv = None
        if intervals is not None and len(intervals):
            iv = np.ascontiguousarray(np.asarray(intervals, dtype=np.uint64).reshape(-1, 2))
        return int(text_begin), int(text_len), chrom_cum, iv

    def compute_mappability(self, params, text_begin=0, text_len=None, chrom_cum_lengths=None, intervals=None,
                            count_fetches=False, return_stats=False):
        """computeMappability<E>(index, text, c, ...) (src/algo.hpp:405-483) for one FASTA file;
        returns the frequency vector c (uint8 / uint16, one value per text position) in host memory."""
        tb, tl, cum, iv = self._file_args(text_begin, text_len, chrom_cum_lengths, intervals)
        p = GmbParams(params.length, params.errors, int(params.rev_compl), int(params.exclude_pseudo),
                      params.value_bits, int(count_fetches), params.block_kmers)
        out = np.zeros(tl, dtype=np.uint16 if params.value_bits == 16 else np.uint8)
        st = GmbMapStats()
        check(_lib.lib().gmb_map_frequencies(self._h, ctypes.byref(p), tb, tl, _ptr(cum), len(cum) - 1, _ptr(iv),
                                             0 if iv is None else len(iv), _ptr(self.seq_to_file),
                                             0 if self.seq_to_file is None else len(self.seq_to_file), _ptr(out),
                                             ctypes.byref(st)))
        return (out, st) if return_stats else out

    def compute_mappability_range(self, params, pos_begin, pos_end, out=None, text_begin=0, text_len=None,
                                  chrom_cum_lengths=None, intervals=None, return_stats=False):
        """The slice [pos_begin, pos_end) of c into host memory (`out`: a numpy array of that length,
        e.g. a view of pinned memory)."""
        tb, tl, cum, iv = self._file_args(text_begin, text_len, chrom_cum_lengths, intervals)
        p = GmbParams(params.length, params.errors, int(params.rev_compl), int(params.exclude_pseudo),
                      params.value_bits, 0, params.block_kmers)
        dt = np.uint16 if params.value_bits == 16 else np.uint8
        if out is None:
            out = np.zeros(int(pos_end) - int(pos_begin), dtype=dt)
        assert out.dtype == dt and out.size >= int(pos_end) - int(pos_begin) and out.flags["C_CONTIGUOUS"]
        st = GmbMapStats()
        check(_lib.lib().gmb_map_frequencies_range(
            self._h, ctypes.byref(p), tb, tl, _ptr(cum), len(cum) - 1, _ptr(iv), 0 if iv is None else len(iv),
            _ptr(self.seq_to_file), 0 if self.seq_to_file is None else len(self.seq_to_file), int(pos_begin),
            int(pos_end), _ptr(out), ctypes.byref(st)))
        return (out, st) if return_stats else out

    def compute_locations(self, params, pos_begin=0, pos_end=None, text_begin=0, text_len=None, chrom_cum_lengths=None,
                          intervals=None, max_locations=0):
        """The csvComputation branch (src/algo.hpp:311-346) for the file-local positions [pos_begin, pos_end):
        -> (offsets uint64[2n+1], loc uint32[m, 2]) with list (j, strand) = loc[offsets[2(j-pos_begin)+strand] :
        offsets[2(j-pos_begin)+strand+1]], rows = (sequence number, offset), sorted; strand 0 = +, 1 = -."""
        tb, tl, cum, iv = self._file_args(text_begin, text_len, chrom_cum_lengths, intervals)
        pos_end = tl if pos_end is None else int(pos_end)
        p = GmbParams(params.length, params.errors, int(params.rev_compl), 0, 16, 0, 1)
        offs, locs, b = [], [], int(pos_begin)
        while b < pos_end:
            r = GmbLocations()
            check(_lib.lib().gmb_map_locations(self._h, ctypes.byref(p), tb, tl, _ptr(cum), len(cum) - 1, _ptr(iv),
                                               0 if iv is None else len(iv), b, pos_end, int(max_locations), ctypes.byref(r)))
            n = int(r.pos_end - r.pos_begin)
            o = np.ctypeslib.as_array(r.offsets, shape=(2 * n + 1,)).copy()
            l = (np.frombuffer(ctypes.string_at(r.loc, 8 * int(r.n_locations)), dtype=np.uint32).reshape(-1, 2).copy()
                 if r.n_locations else np.zeros((0, 2), np.uint32))
            base = offs[-1][-1] if offs else np.uint64(0)
            offs.append(o[(1 if offs else 0):] + base)
            locs.append(l)
            b = int(r.pos_end)
            _lib.lib().gmb_locations_free(ctypes.byref(r))
        return np.concatenate(offs), np.concatenate(locs)

    def compute_runs(self, params, pos_begin=0, pos_end=None, text_begin=0, text_len=None, chrom_cum_lengths=None,
                     intervals=None, return_timings=False):
        """The frequency vector of [pos_begin, pos_end) run-length encoded on the device (what saveWig / saveBedGraph
        scan for, src/output.hpp:73-187) -> (start uint64[n], value uint16[n]); run r = [start[r], start[r+1])."""
        tb, tl, cum, iv = self._file_args(text_begin, text_len, chrom_cum_lengths, intervals)
        pos_end = tl if pos_end is None else int(pos_end)
        p = GmbParams(params.length, params.errors, int(params.rev_compl), int(params.exclude_pseudo),
                      params.value_bits, 0, params.block_kmers)
        r = GmbRuns()
        check(_lib.lib().gmb_map_runs(self._h, ctypes.byref(p), tb, tl, _ptr(cum), len(cum) - 1, _ptr(iv),
                                      0 if iv is None else len(iv), _ptr(self.seq_to_file),
                                      0 if self.seq_to_file is None else len(self.seq_to_file), int(pos_begin), pos_end,
                                      ctypes.byref(r), None))
        n = int(r.n_runs)
        start = np.ctypeslib.as_array(r.start, shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        value = np.ctypeslib.as_array(r.value, shape=(n,)).copy() if n else np.zeros(0, np.uint16)
        tm = (r.kernel_ms, r.rle_ms)
        _lib.lib().gmb_runs_free(ctypes.byref(r))
        return (start, value, tm) if return_timings else (start, value)

    def set_jump_depth(self, depth):
        """-1 = automatic, 0 = no jump tables, 1..16 = maximum table depth (see gmb_index_set_jump_depth)."""
        check(_lib.lib().gmb_index_set_jump_depth(self._h, int(depth)))

    def set_plan_text_size(self, n_symbols):
        """Plan the searches as if the text had n_symbols symbols (0 = the index's own size); counts never change."""
        check(_lib.lib().gmb_index_set_plan_text_size(self._h, int(n_symbols)))

    def progress(self):
        """(positions handed out so far, positions of the call) of the map call in flight on this handle."""
        d, t = ctypes.c_uint64(), ctypes.c_uint64()
        check(_lib.lib().gmb_progress(self._h, ctypes.byref(d), ctypes.byref(t)))
        return int(d.value), int(t.value)

    def refresh_info(self):
        check(_lib.lib().gmb_index_get_info(self._h, ctypes.byref(self.info)))
        return self.info

    def export_bwt(self, rev=False):
        """BWT of T (or of T' if rev) as bytes: 0 = sentinel, 1..4 = A,C,G,T (diagnostics / tests)."""
        out = np.zeros(int(self.info.n_bwt), dtype=np.uint8)
        check(_lib.lib().gmb_index_export_bwt(self._h, int(rev), _ptr(out)))
        return out

    def export_sa(self):
        """Full suffix array of the sentinel-separated text (needs an index built with with_sa=True)."""
        out = np.zeros(int(self.info.n_bwt), dtype=np.uint32)
        check(_lib.lib().gmb_index_export_sa(self._h, _ptr(out)))
        return out

    def compute_mappability_device(self, params, out_ptr, text_begin=0, text_len=None, chrom_cum_lengths=None,
                                   intervals=None, pos_begin=0, pos_end=None, stream=0, count_fetches=False,
                                   sync=True):
        """Same, for the file-local position range [pos_begin, pos_end) only, writing into device memory at
        `out_ptr` (text_len elements, zero-filled by the caller) on CUDA stream `stream`."""
        tb, tl, cum, iv = self._file_args(text_begin, text_len, chrom_cum_lengths, intervals)
        p = GmbParams(params.length, params.errors, int(params.rev_compl), int(params.exclude_pseudo),
                      params.value_bits, int(count_fetches), params.block_kmers)
        st = GmbMapStats()
        check(_lib.lib().gmb_map_frequencies_device(
            self._h, ctypes.byref(p), tb, tl, _ptr(cum), len(cum) - 1, _ptr(iv), 0 if iv is None else len(iv),
            _ptr(self.seq_to_file), 0 if self.seq_to_file is None else len(self.seq_to_file), int(pos_begin),
            tl if pos_end is None else int(pos_end), ctypes.c_void_p(out_ptr), ctypes.c_void_p(stream),
            ctypes.byref(st) if sync else None))
        return st if sync else None


def compute_mappability(index, K, E=0, rev_compl=True, value_bits=16, **kw):
    return index.compute_mappability(SearchParams(K, E, rev_compl, False, value_bits), **kw)
