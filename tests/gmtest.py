"""Test-side helpers (TEST INFRASTRUCTURE): FASTA parsing, synthetic genomes, the ctypes binding of
the CPU oracle (oracle/gm_oracle.c) and a runner for the unmodified reference binary
(oracle/_ref/genmap_ref).  Nothing here is imported by the product package."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libgm_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "genmap_ref")

_CODE = np.full(256, 4, dtype=np.uint8)  # everything that is not ACGT(U) is N (src/indexing.hpp:13-20)
for _i, _ch in enumerate("ACGT"):
    _CODE[ord(_ch)] = _i
    _CODE[ord(_ch.lower())] = _i
_CODE[ord("U")] = 3
_CODE[ord("u")] = 3


def read_fasta(path):
    """-> list of (id, codes uint8[len]); empty records skipped (src/indexing.hpp:228-231)."""
    recs, name, chunks = [], None, []
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None and sum(len(c) for c in chunks):
                    recs.append((name, np.concatenate(chunks)))
                name, chunks = line[1:].decode(), []
            elif line:
                chunks.append(_CODE[np.frombuffer(line, dtype=np.uint8)])
    if name is not None and sum(len(c) for c in chunks):
        recs.append((name, np.concatenate(chunks)))
    # ids cut at the first whitespace if still unique (src/indexing.hpp:238-266)
    short = [n.split()[0] if n.split() else "" for n, _ in recs]
    if len(set(short)) == len(short):
        recs = [(s, c) for s, (_, c) in zip(short, recs)]
    return recs


def write_fasta(path, seqs, names=None, width=80):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s\n" % (names[i] if names else "chr%d" % (i + 1)))
            txt = np.frombuffer(b"ACGTN", dtype=np.uint8)[s].tobytes().decode()
            for j in range(0, len(txt), width):
                f.write(txt[j:j + width] + "\n")


def synth(total, nchr, seed, rep_frac=0.05, mut=0.02):
    """The frozen generator of BASELINE.md §2 -> list of uint8 code arrays."""
    rng = np.random.default_rng(seed)
    per = total // nchr
    out = []
    for _ in range(nchr):
        a = rng.integers(0, 4, per, dtype=np.uint8)
        for _ in range(int(per * rep_frac / 1000)):
            L = int(rng.integers(200, 2000)); src = int(rng.integers(0, per - L)); dst = int(rng.integers(0, per - L))
            seg = a[src:src + L].copy()
            if rng.random() < 0.5:
                seg = 3 - seg[::-1]
            m = rng.random(L) < mut
            seg[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
            a[dst:dst + L] = seg
        out.append(a)
    return out


def repeat_rich(seed, nchr, length, rep_frac=0.3, mut=0.03, with_n=False):
    """Small repeat-heavy genome (many counts > 1, both strands) for E>0 parity cases."""
    rng = np.random.default_rng(seed)
    pool = np.zeros(0, dtype=np.uint8)
    out = []
    for _ in range(nchr):
        parts, n = [], 0
        while n < length:
            if len(pool) > 250 and rng.random() < rep_frac:
                L = int(rng.integers(20, 200)); st = int(rng.integers(0, len(pool) - L))
                seg = pool[st:st + L].copy()
                if rng.random() < 0.5:
                    seg = np.where(seg < 4, 3 - seg, seg)[::-1].astype(np.uint8)
                m = rng.random(L) < mut
                seg[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
            else:
                seg = rng.integers(0, 4, int(rng.integers(50, 300)), dtype=np.uint8)
            parts.append(seg); n += len(seg)
        s = np.concatenate(parts)[:length].copy()
        if with_n:
            for _ in range(max(1, length // 500)):
                s[int(rng.integers(0, length - 3))] = 4
        out.append(s)
        pool = np.concatenate([pool, s])
    return out


def concat(seqs):
    codes = np.concatenate(seqs).astype(np.uint8) if seqs else np.zeros(0, np.uint8)
    limits = np.zeros(len(seqs) + 1, dtype=np.uint64)
    limits[1:] = np.cumsum([len(s) for s in seqs])
    return np.ascontiguousarray(codes), limits


# ---------------------------------------------------------------------------------------------
# oracle binding
# ---------------------------------------------------------------------------------------------
class _Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in
                ("K", "E", "revcompl", "exclude_pseudo", "value_bits", "infix_len", "threads", "copy_shortcut")]


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return ORACLE_SO


_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = ctypes.CDLL(ORACLE_SO)
        vp, u64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32
        L.gmo_index_build.restype = vp
        L.gmo_index_build.argtypes = [vp, vp, u32]
        L.gmo_index_from_bwt.restype = vp
        L.gmo_index_from_bwt.argtypes = [vp, vp, u32, vp, vp, u32, vp]
        L.gmo_index_free.argtypes = [vp]
        L.gmo_index_bwt_len.restype = u64
        L.gmo_index_bwt_len.argtypes = [vp]
        L.gmo_index_sigma.restype = u32
        L.gmo_index_sigma.argtypes = [vp]
        L.gmo_index_get_bwt.argtypes = [vp, ctypes.c_int, vp]
        L.gmo_index_get_sa.argtypes = [vp, vp]
        L.gmo_map.restype = ctypes.c_int
        L.gmo_map.argtypes = [vp, ctypes.POINTER(_Params), u64, u64, vp, u32, vp, u64, vp, vp]
        L.gmo_brute.restype = ctypes.c_int
        L.gmo_brute.argtypes = [vp, vp, u32, ctypes.POINTER(_Params), u64, u64, vp, u32, vp, u64, vp, vp]
        L.gmo_brute_locations.restype = u64
        L.gmo_brute_locations.argtypes = [vp, vp, u32, u32, u32, u64, ctypes.c_int, vp, vp, u64]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _params(K, E, revcompl, exclude_pseudo, value_bits, infix_len, threads, copy_shortcut):
    return _Params(K, E, int(revcompl), int(exclude_pseudo), value_bits, infix_len, threads, int(copy_shortcut))


def _prep(limits, seq_to_file, file_no, intervals):
    """file-level arguments shared by every implementation: (text_begin, text_len, chrom_cum)."""
    n_seq = len(limits) - 1
    if seq_to_file is None:
        seq_to_file = np.zeros(n_seq, dtype=np.uint32)
    seq_to_file = np.ascontiguousarray(seq_to_file, dtype=np.uint32)
    sel = np.nonzero(seq_to_file == file_no)[0]
    s0, s1 = int(sel[0]), int(sel[-1]) + 1
    text_begin, text_len = int(limits[s0]), int(limits[s1] - limits[s0])
    chrom_cum = np.ascontiguousarray(limits[s0:s1 + 1] - limits[s0], dtype=np.uint64)
    iv = None
    if intervals is not None and len(intervals):
        iv = np.ascontiguousarray(np.asarray(intervals, dtype=np.uint64).reshape(-1, 2))
    return seq_to_file, text_begin, text_len, chrom_cum, iv


class Oracle:
    """CPU restatement of the reference path (oracle/gm_oracle.c)."""

    def __init__(self, seqs, seq_to_file=None, bwt=None):
        self.L = oracle_lib()
        self.codes, self.limits = concat(seqs)
        self.n_seq = len(seqs)
        self.seq_to_file = seq_to_file
        if bwt is None:
            self.h = self.L.gmo_index_build(_ptr(self.codes), _ptr(self.limits), self.n_seq)
        else:
            f, r, sigma = bwt
            self.h = self.L.gmo_index_from_bwt(_ptr(self.codes), _ptr(self.limits), self.n_seq,
                                               _ptr(np.ascontiguousarray(f)), _ptr(np.ascontiguousarray(r)), sigma, None)
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.gmo_index_free(self.h)
            self.h = None

    @property
    def sigma(self):
        return self.L.gmo_index_sigma(self.h)

    def bwt(self, rev=False):
        out = np.zeros(self.L.gmo_index_bwt_len(self.h), dtype=np.uint8)
        self.L.gmo_index_get_bwt(self.h, int(rev), _ptr(out))
        return out

    def sa(self):
        out = np.zeros(self.L.gmo_index_bwt_len(self.h), dtype=np.uint64)
        assert self.L.gmo_index_get_sa(self.h, _ptr(out)) == 0
        return out

    def map(self, K, E, revcompl=True, exclude_pseudo=False, value_bits=16, file_no=0, intervals=None,
            infix_len=0, threads=0, copy_shortcut=False):
        stf, tb, tl, cum, iv = _prep(self.limits, self.seq_to_file, file_no, intervals)
        out = np.zeros(tl, dtype=np.uint16 if value_bits == 16 else np.uint8)
        p = _params(K, E, revcompl, exclude_pseudo, value_bits, infix_len, threads, copy_shortcut)
        rc = self.L.gmo_map(self.h, ctypes.byref(p), tb, tl, _ptr(cum), len(cum) - 1, _ptr(iv),
                            0 if iv is None else len(iv), _ptr(stf), _ptr(out))
        if rc != 0:
            raise RuntimeError("gmo_map failed: %d" % rc)
        return out


def brute(seqs, K, E, revcompl=True, exclude_pseudo=False, value_bits=16, seq_to_file=None, file_no=0,
          intervals=None):
    L = oracle_lib()
    codes, limits = concat(seqs)
    stf, tb, tl, cum, iv = _prep(limits, seq_to_file, file_no, intervals)
    out = np.zeros(tl, dtype=np.uint16 if value_bits == 16 else np.uint8)
    p = _params(K, E, revcompl, exclude_pseudo, value_bits, 0, 0, False)
    rc = L.gmo_brute(_ptr(codes), _ptr(limits), len(seqs), ctypes.byref(p), tb, tl, _ptr(cum), len(cum) - 1,
                     _ptr(iv), 0 if iv is None else len(iv), _ptr(stf), _ptr(out))
    assert rc == 0
    return out


def brute_locations(seqs, K, E, positions, revcompl=True):
    """Definition-level csv lists: {concatenated-text position: (plus, minus)} with plus/minus = lists of
    (sequence, offset) in sorted order (oracle/gm_oracle.c:gmo_brute_locations)."""
    L = oracle_lib()
    codes, limits = concat(seqs)
    cap = int(limits[-1]) + 1
    sq, ps = np.zeros(cap, dtype=np.uint32), np.zeros(cap, dtype=np.uint32)
    out = {}
    for pos in positions:
        both = []
        for strand in ((0, 1) if revcompl else (0,)):
            n = L.gmo_brute_locations(_ptr(codes), _ptr(limits), len(seqs), K, E, int(pos), strand, _ptr(sq), _ptr(ps), cap)
            both.append(list(zip(sq[:n].tolist(), ps[:n].tolist())))
        if not revcompl:
            both.append([])
        out[int(pos)] = tuple(both)
    return out


def valid_starts(limits, K, text_begin=0, text_len=None):
    """file-local positions whose K-window stays inside its sequence (the k-mers the reference emits, src/algo.hpp:380)"""
    limits = np.asarray(limits, dtype=np.int64)
    end = int(limits[-1]) if text_len is None else text_begin + text_len
    out = []
    for s in range(len(limits) - 1):
        b, e = int(limits[s]), int(limits[s + 1])
        if b >= text_begin and e <= end:
            out.extend(range(b - text_begin, max(b, e - K + 1) - text_begin))
    return out


def csv_render(lists, chrom_cum, file_names, file_last_seq, revcompl):
    """saveCsv (src/output.hpp:189-288) for one FASTA file from {file-local position: (plus, minus)} lists of
    (global sequence, offset): one line per k-mer with at least one occurrence (src/algo.hpp:377-385)."""
    cum = np.asarray(chrom_cum, dtype=np.int64)
    lines = ['"k-mer"' + "".join(';"+ strand %s"' % f for f in file_names)
             + ("".join(';"- strand %s"' % f for f in file_names) if revcompl else "")]
    for j in sorted(lists):
        plus, minus = lists[j]
        if not plus and not minus:
            continue
        ch = int(np.searchsorted(cum, j, side="right") - 1)
        line = "%d,%d" % (ch, j - cum[ch])
        for lst in ((plus, minus) if revcompl else (plus,)):
            i, before = 0, 0
            for last in file_last_seq:
                col = []
                while i < len(lst) and lst[i][0] <= last:
                    col.append("%d,%d" % (lst[i][0] - before, lst[i][1]))
                    i += 1
                line += ";" + "|".join(col)
                before = last + 1
        lines.append(line)
    return "\n".join(lines) + "\n"


def lists_from_arrays(offsets, loc, pos_begin=0):
    """(offsets, loc) of gmb_map_locations / Index.compute_locations -> {position: (plus, minus)}"""
    loc = [tuple(x) for x in np.asarray(loc).tolist()]
    off = np.asarray(offsets).astype(np.int64)
    return {pos_begin + j: (loc[off[2 * j]:off[2 * j + 1]], loc[off[2 * j + 1]:off[2 * j + 2]]) for j in range((len(off) - 1) // 2)}


# ---------------------------------------------------------------------------------------------
# the unmodified reference binary (oracle/_ref/genmap_ref), when present
# ---------------------------------------------------------------------------------------------
def have_reference():
    return os.access(REF_BIN, os.X_OK)


def run_reference(fasta_or_dir, K, E, flags=(), value_bits=16, threads=None, selection=None, verbose_time=False):
    """index + map with the reference binary -> {basename: np.ndarray}; optional (dict, seconds)."""
    with tempfile.TemporaryDirectory() as tmp:
        idx, out = os.path.join(tmp, "index"), os.path.join(tmp, "out")
        os.mkdir(out)
        flag = "-FD" if os.path.isdir(fasta_or_dir) else "-F"
        subprocess.run([REF_BIN, "index", flag, fasta_or_dir, "-I", idx], check=True, stdout=subprocess.DEVNULL)
        cmd = [REF_BIN, "map", "-I", idx, "-O", out, "-K", str(K), "-E", str(E), "-r",
               "-fl" if value_bits == 16 else "-fs", "-v"] + list(flags)
        if threads:
            cmd += ["-T", str(threads)]
        if selection:
            cmd += ["-S", selection]
        res = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True)
        secs = None
        for line in res.stdout.replace("\r", "\n").split("\n"):
            if line.startswith("Mappability computed in"):
                secs = float(line.split()[3])
        ext = ".freq16" if value_bits == 16 else ".freq8"
        dt = np.uint16 if value_bits == 16 else np.uint8
        outs = {}
        for fn in sorted(os.listdir(out)):
            if fn.endswith(ext):
                outs[fn[:-len(".genmap" + ext)]] = np.fromfile(os.path.join(out, fn), dtype=dt)
        return (outs, secs) if verbose_time else outs


# ---------------------------------------------------------------------------------------------
# the reference's golden cases (tests/golden/reference_cases, copied from tests/test_cases)
# flags from tests/CMakeLists.txt:56-73
# ---------------------------------------------------------------------------------------------
CASES = {
    "1a": dict(dir=False, K=3, E=0, rc=False, ep=False), "1b": dict(dir=False, K=3, E=0, rc=True, ep=False),
    "1c": dict(dir=False, K=3, E=0, rc=False, ep=False), "1d": dict(dir=False, K=3, E=0, rc=True, ep=False),
    "1e": dict(dir=False, K=3, E=1, rc=False, ep=False), "1f": dict(dir=False, K=3, E=1, rc=True, ep=False),
    "1g": dict(dir=False, K=3, E=1, rc=True, ep=False),
    "2a": dict(dir=False, K=4, E=0, rc=False, ep=False), "2b": dict(dir=False, K=4, E=0, rc=True, ep=False),
    "2c": dict(dir=False, K=4, E=0, rc=False, ep=False), "2d": dict(dir=False, K=4, E=0, rc=True, ep=False),
    "2e": dict(dir=False, K=4, E=0, rc=True, ep=False),
    "3a": dict(dir=True, K=4, E=0, rc=False, ep=False), "3b": dict(dir=True, K=4, E=0, rc=True, ep=False),
    "3c": dict(dir=True, K=4, E=0, rc=False, ep=True), "3d": dict(dir=True, K=4, E=0, rc=True, ep=True),
    "3e": dict(dir=True, K=4, E=0, rc=True, ep=True), "3f": dict(dir=True, K=4, E=0, rc=True, ep=True),
}


def load_case(case):
    """-> (files: [(basename_without_ext, [(id, codes)])], selection: {seqid: [(b,e)]} or None, folder)"""
    folder = os.path.join(GOLDEN, "reference_cases", "case_" + case)
    fas = sorted(f for f in os.listdir(folder) if f.endswith(".fa"))  # -FD sorts by file name (src/indexing.hpp:407)
    files = [(f[:-3], read_fasta(os.path.join(folder, f))) for f in fas]
    sel = None
    bed = os.path.join(folder, "subset.bed")
    if os.path.exists(bed):
        sel = {}
        for line in open(bed):
            p = line.split()
            if len(p) >= 3:
                sel.setdefault(p[0], []).append((int(p[1]), int(p[2])))
    return files, sel, folder


def case_layout(files):
    """-> seqs (all files, index order), seq_to_file, per-file list of sequence names"""
    seqs, stf, names = [], [], []
    for fi, (_, recs) in enumerate(files):
        names.append([n for n, _ in recs])
        for _, c in recs:
            seqs.append(c); stf.append(fi)
    return seqs, np.asarray(stf, dtype=np.uint32), names


def file_intervals(sel, recs):
    """BED selection -> file-local cumulative intervals (src/mappability.hpp:334-358); None if the
    file has no interval (then the reference writes no output for it, :309)."""
    if sel is None:
        return []
    iv, cum = [], 0
    for name, codes in recs:
        for b, e in sel.get(name, []):
            iv.append((cum + b, cum + e))
        cum += len(codes)
    return iv if iv else None


# ---------------------------------------------------------------------------------------------
# host simulation of the kernel's state machine (tests/hostsim) — CPU debugging aid, tests only
# ---------------------------------------------------------------------------------------------
HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
HOSTSIM_SO = os.path.join(HOSTSIM_DIR, "_build_hostsim.so")
_hs = None


def hostsim_lib():
    global _hs
    if _hs is None:
        csrc = os.path.join(ROOT, "genmap_b200", "csrc")
        srcs = [os.path.join(HOSTSIM_DIR, "hostsim.cpp"), os.path.join(csrc, "gmb_host.cpp")]
        deps = srcs + [os.path.join(csrc, f) for f in ("gmb_core.h", "gmb_layout.h", "gmb_host.h", "sais.hpp")]
        if not os.path.exists(HOSTSIM_SO) or any(os.path.getmtime(d) > os.path.getmtime(HOSTSIM_SO) for d in deps):
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                            "-o", HOSTSIM_SO] + srcs, check=True)
        L = ctypes.CDLL(HOSTSIM_SO)
        vp, u64, u32, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
        L.hs_build.restype = ci
        L.hs_build.argtypes = [vp, vp, u32, ci, ctypes.POINTER(vp), ctypes.POINTER(u64)]
        L.hs_free.argtypes = [vp]
        L.hs_export_bwt.argtypes = [vp, ci, vp]
        L.hs_export_sa.restype = ci
        L.hs_export_sa.argtypes = [vp, vp]
        L.hs_has_n_selftest.restype = ci
        L.hs_has_n_selftest.argtypes = [u64]
        L.hs_step_tables.restype = ci
        L.hs_step_tables.argtypes = [u32, u32, ctypes.POINTER(u32), vp]
        L.hs_map.restype = ci
        L.hs_map.argtypes = [vp, u32, u32, ci, ci, u64, u64, vp, u32, vp, u64, u64, u64, vp, vp,
                             ci, ctypes.POINTER(ctypes.c_ulonglong), vp, u32, u32]
        L.hs_locate.restype = ci
        L.hs_locate.argtypes = [vp, u32, u32, ci, u64, u64, vp, u32, vp, u64, u64, u64, ci, vp, ctypes.POINTER(vp)]
        _hs = L
    return _hs


class HostSim:
    def __init__(self, seqs, with_sa=False):
        self.L = hostsim_lib()
        self.codes, self.limits = concat(seqs)
        self.n_seq = len(seqs)
        blob, nbytes = ctypes.c_void_p(), ctypes.c_uint64()
        rc = self.L.hs_build(_ptr(self.codes), _ptr(self.limits), self.n_seq, int(with_sa), ctypes.byref(blob), ctypes.byref(nbytes))
        if rc != 0:
            raise RuntimeError("hs_build failed")
        self.blob, self.nbytes = blob, nbytes.value
        self.n_bwt = len(self.codes) + self.n_seq

    def __del__(self):
        if getattr(self, "blob", None):
            self.L.hs_free(self.blob)
            self.blob = None

    def blob_bytes(self):
        return ctypes.string_at(self.blob, self.nbytes)

    def bwt(self, rev=False):
        out = np.zeros(self.n_bwt, dtype=np.uint8)
        self.L.hs_export_bwt(self.blob, int(rev), _ptr(out))
        return out

    def sa(self):
        out = np.zeros(self.n_bwt, dtype=np.uint32)
        assert self.L.hs_export_sa(self.blob, _ptr(out)) == 0
        return out

    def map(self, K, E, revcompl=True, value_bits=16, seq_to_file=None, file_no=0, intervals=None,
            pos_begin=0, pos_end=None, return_fetches=False, jump_depth=-1, exclude_pseudo=False, block_kmers=0):
        stf, tb, tl, cum, iv = _prep(self.limits, seq_to_file, file_no, intervals)
        out = np.zeros(tl, dtype=np.uint16 if value_bits == 16 else np.uint8)
        f, lr = (ctypes.c_ulonglong * 13)(), ctypes.c_ulonglong(0)
        rc = self.L.hs_map(self.blob, K, E, int(revcompl), value_bits, tb, tl, _ptr(cum), len(cum) - 1, _ptr(iv),
                           0 if iv is None else len(iv), pos_begin, tl if pos_end is None else pos_end, _ptr(out),
                           f, jump_depth, ctypes.byref(lr), _ptr(stf) if exclude_pseudo else None, file_no,
                           block_kmers)
        if rc != 0:
            raise RuntimeError("hs_map failed: %d" % rc)
        self.last_lut_reads = lr.value
        self.last_fetch_stats = list(f)  # total, by interval size [8], thin paths, iterations, located entries, text reads
        return (out, f[0]) if return_fetches else out

    def locate(self, K, E, revcompl=True, seq_to_file=None, file_no=0, intervals=None, pos_begin=0, pos_end=None,
               jump_depth=-1):
        """csv lists of the file-local positions [pos_begin, pos_end) -> {position: (plus, minus)}, (seq, offset) pairs"""
        stf, tb, tl, cum, iv = _prep(self.limits, seq_to_file, file_no, intervals)
        pos_end = tl if pos_end is None else pos_end
        off = np.zeros(2 * (pos_end - pos_begin) + 1, dtype=np.uint64)
        rows = ctypes.c_void_p()
        rc = self.L.hs_locate(self.blob, K, E, int(revcompl), tb, tl, _ptr(cum), len(cum) - 1, _ptr(iv),
                              0 if iv is None else len(iv), pos_begin, pos_end, jump_depth, _ptr(off), ctypes.byref(rows))
        if rc != 0:
            raise RuntimeError("hs_locate failed: %d" % rc)
        n = int(off[-1])
        r = np.frombuffer(ctypes.string_at(rows, 4 * n), dtype=np.uint32).astype(np.int64) if n else np.zeros(0, np.int64)
        self.L.hs_free(rows)
        seq_start = np.asarray(self.limits, dtype=np.int64) + np.arange(self.n_seq + 1)
        sq = np.searchsorted(seq_start, r, side="right") - 1
        pairs = list(zip(sq.tolist(), (r - seq_start[sq]).tolist()))
        return {pos_begin + j: (pairs[int(off[2 * j]):int(off[2 * j + 1])], pairs[int(off[2 * j + 1]):int(off[2 * j + 2])])
                for j in range(pos_end - pos_begin)}


# ---------------------------------------------------------------------------------------------
# writing the reference's on-disk index from a BWT + SA computed elsewhere (oracle/seqan_index.c)
# ---------------------------------------------------------------------------------------------
def write_seqan_index(directory, files, bwt_fwd, bwt_rev, sa, sampling=10):
    """files: [(file_name, [(seq_name, codes)])] in index order.  bwt_*: uint8 rows, 0 = sentinel, 1..4 = ACGT.
    sa: uint32 suffix array of the sentinel-separated text.  Writes everything `genmap_ref map` opens."""
    L = oracle_lib()
    vp, u64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32
    L.gmo_seqan_write_lf.restype = ctypes.c_int
    L.gmo_seqan_write_lf.argtypes = [ctypes.c_char_p, vp, u64, vp]
    L.gmo_seqan_write_sa.restype = ctypes.c_int
    L.gmo_seqan_write_sa.argtypes = [ctypes.c_char_p, vp, u64, vp, u32, u32]
    L.gmo_seqan_write_packed_text.restype = ctypes.c_int
    L.gmo_seqan_write_packed_text.argtypes = [ctypes.c_char_p, vp, u64]
    os.makedirs(directory, exist_ok=True)
    base = os.path.join(directory, "index")
    seqs = [c for _, recs in files for _, c in recs]
    codes, limits = concat(seqs)
    n_seq, n = len(seqs), int(limits[-1]) + len(seqs)

    def string_set(path, strings):
        raw = [s.encode() for s in strings]
        with open(path + ".concat", "wb") as f:
            f.write(b"".join(raw))
        np.concatenate([[0], np.cumsum([len(r) for r in raw])]).astype("<u8").tofile(path + ".limits")

    directory_flag = "true" if len(files) > 1 else "false"
    string_set(base + ".info", ["alphabet_size:4", "sa_dimensions_i1:16", "sa_dimensions_i2:32", "bwt_dimensions:32",
                                "sampling_rate:%d" % sampling, "fasta_directory:" + directory_flag, "packed_text:true"])
    string_set(base + ".ids", ["%s;%d;%s" % (fn, len(c), name) for fn, recs in files for name, c in recs])
    limits.astype("<u8").tofile(base + ".txt.limits")
    assert L.gmo_seqan_write_packed_text((base + ".txt.concat").encode(), _ptr(codes), len(codes)) == 0
    for rev, bwt in ((False, bwt_fwd), (True, bwt_rev)):
        pre = base + (".rev.lf" if rev else ".lf")
        bwt = np.ascontiguousarray(bwt, dtype=np.uint8)
        counts = np.zeros(4, dtype=np.uint64)
        assert L.gmo_seqan_write_lf(pre.encode(), _ptr(bwt), n, _ptr(counts)) == 0
        pst = np.zeros(5, dtype="<u4")
        pst[0] = n_seq
        pst[1:] = n_seq + np.cumsum(counts)
        pst.tofile(pre + ".pst")
        with open(pre + ".drs", "wb") as f:
            f.write(b"\x00")
    seq_start = np.ascontiguousarray(limits + np.arange(n_seq + 1, dtype=np.uint64))
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    assert L.gmo_seqan_write_sa((base + ".sa").encode(), _ptr(sa), n, _ptr(seq_start), n_seq, sampling) == 0
    np.array([n], dtype="<u8").tofile(base + ".sa.len")
    return directory


def fragmented_genome(seed, total=12000, with_n=False):
    """Many short sequences (10 - 400 bases) with repeats shared between them and reverse-complemented copies: windows,
    table-key contexts and candidate alignments cross sequence boundaries everywhere (what verify_located must reject)."""
    rng = np.random.default_rng(seed)
    pool = rng.integers(0, 4, 3000, dtype=np.uint8)
    seqs, n = [], 0
    while n < total:
        L = int(rng.integers(10, 400))
        if rng.random() < 0.5:
            st = int(rng.integers(0, len(pool) - L))
            s = pool[st:st + L].copy()
            if rng.random() < 0.4:
                s = (3 - s)[::-1].copy()
            m = rng.random(L) < 0.02
            s[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
        else:
            s = rng.integers(0, 4, L, dtype=np.uint8)
        if with_n and rng.random() < 0.1:
            s[int(rng.integers(0, L))] = 4
        seqs.append(s); n += L
    return seqs
