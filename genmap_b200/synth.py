"""The frozen synthetic-genome generator of BASELINE.md §2 (iid ACGT + planted, mutated, half
reverse-complemented repeats).  Data generation only; not part of the hot path."""
import numpy as np


def synth_genome(total, nchr, seed, rep_frac=0.05, mut=0.02):
    """-> list of nchr uint8 code arrays (0..3 = ACGT), total // nchr bases each."""
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


def synth_pangenome(file_bp, n_files, nchr=3, seed=46):
    """BASELINE config 5: n_files FASTA files of nchr sequences; file g = the base genome (frozen generator, `seed`) with
    g % substitutions.  -> (list of n_files * nchr code arrays in index order, seq_to_file uint32)."""
    base = synth_genome(int(file_bp), nchr, seed)
    rng = np.random.default_rng(seed + 1)
    seqs, stf = [], []
    for g in range(n_files):
        for s in base:
            s = s.copy()
            if g:
                n_sub = int(len(s) * 0.01 * g)
                idx = rng.choice(len(s), n_sub, replace=False) if len(s) < 5_000_000 else np.unique(rng.integers(0, len(s), n_sub))
                s[idx] = (s[idx] + rng.integers(1, 4, len(idx), dtype=np.uint8)) & 3  # a substitution always changes the base
            seqs.append(s); stf.append(g)
    return seqs, np.asarray(stf, dtype=np.uint32)


def write_fasta(path, seqs, width=80):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">chr%d\n" % (i + 1))
            txt = lut[s]
            full = len(txt) // width * width
            if full:
                body = np.concatenate([txt[:full].reshape(-1, width),
                                       np.full((full // width, 1), 10, dtype=np.uint8)], axis=1)
                f.write(body.tobytes())
            if full < len(txt):
                f.write(txt[full:].tobytes() + b"\n")
