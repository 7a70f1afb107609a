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


def synth_pangenome(file_bp, n_files, nchr=3, seed=46, threads=8):
    """BASELINE config 5: n_files FASTA files of nchr sequences; file g = the base genome (frozen generator, `seed`) with
    g % substitutions, every (file, sequence) from its own seeded stream so that the files can be made
    side by side.  -> (list of n_files * nchr code arrays in index order, seq_to_file uint32)."""
    from concurrent.futures import ThreadPoolExecutor
    base = synth_genome(int(file_bp), nchr, seed)

    def one(gc):
        g, c = gc
        s = base[c].copy()
        if g:
            rng = np.random.default_rng([seed + 1, g, c])
            idx = rng.integers(0, len(s), int(len(s) * 0.01 * g))  # (a position drawn twice is substituted once)
            s[idx] = (s[idx] + rng.integers(1, 4, len(idx), dtype=np.uint8)) & 3  # a substitution always changes the base
        return s

    jobs = [(g, c) for g in range(n_files) for c in range(nchr)]
    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        seqs = list(pool.map(one, jobs))
    return seqs, np.asarray([g for g, _ in jobs], dtype=np.uint32)


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
