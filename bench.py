#!/usr/bin/env python
"""bench.py — genome positions/sec of the (K,E)-frequency hot path on a synthetic genome.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--genome-mbp 3000 --nchr 24 --seed 45] [--kmer 30] [--errors 0] [--batch-mpos 256]

A step = one pass of the hot path over one batch of consecutive k-mer start positions.
  value : positions/s, index and output resident in HBM, CUDA events on the launching stream, max over ranks
  e2e   : the same metric through the host-buffer C ABI call gmb_map_frequencies_range (result slice D2H
          into pinned host memory inside the timed region)
  roofline: algorithmic rank-block bytes (counted fetches x 32 B) / kernel time, vs the measured HBM copy peak
  cpu_baseline: the unmodified reference binary (oracle/_ref/genmap_ref; the oracle port only if it is missing) on
          this box's host cores, on a bounded window of the same genome
  parity  : the reference's counts on that window compared with the GPU's (bit-exact or the line says so)
One JSON line on stdout (rank 0); progress on stderr.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genome positions/sec at (K,E) on 3 Gbp"
UNIT = "positions/s"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="genome", choices=["genome", "pangenome"],
                    help="genome: BASELINE configs 2-4 (one synthetic genome); pangenome: config 5 (10 x 300 Mbp files, K=50 E=2 -ep)")
    ap.add_argument("--pan-files", type=int, default=10)
    ap.add_argument("--pan-file-mbp", type=float, default=300.0)
    ap.add_argument("--pan-cpu-mbp", type=float, default=3.0, help="per-file size of the scale model the reference binary is run on")
    ap.add_argument("--genome-mbp", type=float, default=3000.0)
    ap.add_argument("--nchr", type=int, default=24)
    ap.add_argument("--seed", type=int, default=45)
    ap.add_argument("--kmer", type=int, default=30)
    ap.add_argument("--errors", type=int, default=0)
    ap.add_argument("--batch-mpos", type=float, default=0.0, help="batch size in Mi positions (0 = by E)")
    ap.add_argument("--extras", default="1,2", help="other E values measured after the main line ('' = none)")
    ap.add_argument("--jump-depth", type=int, default=-1, help="-1 auto, 0 off, 1..16 max jump-table depth")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work of the cpu_baseline sample")
    ap.add_argument("--ref-step-seconds", type=float, default=4.0, help="--impl reference: CPU work per step")
    return ap.parse_args()


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                sm.append(float(p[1])); mx.append(float(p[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class _DeviceBytes:
    """__cuda_array_interface__ view of raw device memory (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


NUMA_NOTE = "not bound"


def bind_to_gpu_numa_node(local):
    """Run this rank (and so its pinned-buffer allocations: first touch) on the CPUs of the GPU's own NUMA node."""
    global NUMA_NOTE
    try:
        import torch
        props = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            NUMA_NOTE = "single node"
            return
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            NUMA_NOTE = "rank bound to NUMA node %d of its GPU (%d CPUs)" % (node, len(cpus))
    except Exception as ex:
        NUMA_NOTE = "not bound (%s)" % type(ex).__name__


def genome_limits(total, nchr):
    per = total // nchr
    return np.arange(nchr + 1, dtype=np.uint64) * np.uint64(per)


def default_batch(E):
    return {0: 256.0, 1: 64.0, 2: 8.0, 3: 1.0, 4: 0.25}[E]


# --------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation of the path on this box's host cores, on a bounded window of
# the same genome.  Preferred: the UNMODIFIED reference binary (oracle/_ref/genmap_ref) reading an index in
# its own on-disk format, written from the BWT + suffix array the GPU builder produced (its own `genmap
# index` needs ~45 min of single-threaded divsufsort at 3 Gbp).  Fallback: the oracle port (OpenMP).
# --------------------------------------------------------------------------------------------------------
def _window(n_text, per, K, npos):
    start = (n_text // 2 // per) * per + per // 3  # a fixed window inside one chromosome
    b = min(start, n_text - npos - K)
    return b, b + npos


class ReferenceCpu:
    """`genmap_ref map -S window.bed` on a reference-format index of the bench genome."""
    kind = "reference"

    def __init__(self, seqs, ix):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import gmtest as T
        self.T = T
        if not T.have_reference():
            raise RuntimeError("oracle/_ref/genmap_ref is not present")
        self.cores = os.cpu_count() or 1
        self.per = len(seqs[0])
        self.n_text = sum(len(s) for s in seqs)
        t0 = time.time()
        bwt_f, bwt_r, sa = ix.export_bwt(False), ix.export_bwt(True), ix.export_sa()
        tmp_root = os.environ.get("GMB_TMPDIR")
        if not tmp_root and os.path.isdir("/dev/shm"):  # tmpfs: neither the index read nor the 6 GB raw output hits a disk
            st = os.statvfs("/dev/shm")
            if st.f_bavail * st.f_frsize > 8 * self.n_text + (16 << 30):
                tmp_root = "/dev/shm"
        self.dir = tempfile.mkdtemp(prefix="gmb_refidx_", dir=tmp_root)
        files = [("genome.fa", [("chr%d" % (i + 1), s) for i, s in enumerate(seqs)])]
        T.write_seqan_index(os.path.join(self.dir, "index"), files, bwt_f, bwt_r, sa)
        del bwt_f, bwt_r, sa
        log("reference-format index written to %s in %.1f s" % (self.dir, time.time() - t0))

    def run(self, K, E, npos):
        b, e = _window(self.n_text, self.per, K, npos)
        chrom, off = b // self.per, b % self.per
        bed = os.path.join(self.dir, "window.bed")
        with open(bed, "w") as f:
            f.write("chr%d\t%d\t%d\n" % (chrom + 1, off, off + (e - b)))
        out = os.path.join(self.dir, "out")
        os.makedirs(out, exist_ok=True)
        cmd = [self.T.REF_BIN, "map", "-I", os.path.join(self.dir, "index"), "-O", out, "-K", str(K), "-E", str(E),
               "-r", "-fl", "-T", str(self.cores), "-v", "-S", bed]
        res = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True)
        # the counts the reference wrote for this window (saveRaw, src/output.hpp:10-31): kept for the parity check
        raw = np.memmap(os.path.join(out, "genome.genmap.freq16"), dtype=np.uint16, mode="r")
        self.last_window = (b, e, np.array(raw[b:e]))
        del raw
        for line in res.stdout.replace("\r", "\n").split("\n"):
            if line.startswith("Mappability computed in"):
                return max(float(line.split()[3]), 0.005)
        raise RuntimeError("no timing line in the reference's output")

    def sample(self, npos, K):
        b, e = _window(self.n_text, self.per, K, npos)
        return ("`genmap_ref map -S` on %d consecutive positions of chr%d of the same genome (<50%% of the text: "
                "per-position work, copy shortcut off), -T %d; time = its own 'Mappability computed in' line minus the "
                "same line for a 1024-position window (%.2f s: allocation + raw output of the whole-genome vector)"
                % (npos, b // self.per + 1, self.cores, getattr(self, "fixed_s", 0.0)))

    def close(self):
        import shutil
        shutil.rmtree(self.dir, ignore_errors=True)


class PortCpu:
    """The oracle port (oracle/gm_oracle.c, OpenMP) on BWTs exported from the product index."""
    kind = "port"

    def __init__(self, seqs, ix):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import gmtest as T
        self.cores = os.cpu_count() or 1
        t0 = time.time()
        self.orc = T.Oracle(seqs, bwt=(ix.export_bwt(False), ix.export_bwt(True), 4))
        log("oracle rank structure built in %.1f s" % (time.time() - t0))
        self.n_text, self.per = int(self.orc.limits[-1]), int(self.orc.limits[1])

    def run(self, K, E, npos):
        b, e = _window(self.n_text, self.per, K, npos)
        t = time.time()
        c = self.orc.map(K, E, intervals=[(b, e)], threads=self.cores)
        dt = time.time() - t
        self.last_window = (b, e, np.array(c[b:e]))
        return dt

    def sample(self, npos, K):
        b, _ = _window(self.n_text, self.per, K, npos)
        return "oracle port on %d consecutive positions of chr%d (copy shortcut off), %d OpenMP threads" % (
            npos, b // self.per + 1, self.cores)

    def close(self):
        pass


def index_for_cpu_arm(seqs, ix, device):
    """An index of the same genome that holds the suffix array (rebuilt on the GPU: ~2 s at 3 Gbp)."""
    import genmap_b200 as gm
    import torch
    if ix.info.has_sa:
        return ix
    ix.close()
    torch.cuda.empty_cache()
    return gm.Index.build(seqs, device=device, on_gpu=True, with_sa=True)


def make_cpu_arm(seqs, ix):
    try:
        return ReferenceCpu(seqs, ix)
    except Exception as ex:
        log("reference binary arm unavailable (%r): falling back to the oracle port" % (ex,))
        return PortCpu(seqs, ix)


def cpu_rate(arm, K, E, seconds, steps=1, warmup=0):
    """-> (positions/s, per-step ms, positions per step).  Per-call fixed cost (the reference allocates and
    writes the vector of the WHOLE file whatever the window) is measured on a 1024-position window and removed."""
    fixed = 0.0
    if arm.kind == "reference":
        fixed = min(arm.run(K, E, 1024), arm.run(K, E, 1024))
        arm.fixed_s = fixed
    pilot = {0: 2_000_000, 1: 500_000, 2: 50_000, 3: 5_000, 4: 1_000}[E]
    if arm.kind == "port":
        pilot //= 10
    dt = max(arm.run(K, E, pilot) - fixed, 0.02)
    npos = int(max(pilot, min(arm.per // 2, pilot * seconds / dt)))
    times = []
    for i in range(warmup + steps):
        dt = max(arm.run(K, E, npos) - fixed, 0.02)
        if i >= warmup:
            times.append(dt)
    return npos * len(times) / sum(times), [t * 1e3 for t in times], npos


def parity_check(arm, ix, K, E):
    """The counts the CPU arm produced on its last window against the GPU's on the same positions."""
    import genmap_b200 as gm
    b, e, ref = arm.last_window
    got = ix.compute_mappability_range(gm.SearchParams(K, E), b, e)
    diff = np.nonzero(got != ref)[0]
    return {"positions": int(e - b), "window": [int(b), int(e)], "equal": bool(len(diff) == 0), "mismatches": int(len(diff)),
            "against": arm.kind, "nonunique_in_window": int((ref > 1).sum()), "max_count": int(ref.max()) if len(ref) else 0}

# --------------------------------------------------------------------------------------------------------
# BASELINE config 5: multi-FASTA pan-genome, --exclude-pseudo (src/algo.hpp:351-361: distinct FASTA files with an
# occurrence).  All files are indexed together (text + full suffix array in HBM, replicated); a step = one batch
# of positions of one file per GPU, the files' positions range-partitioned over the ranks.
# --------------------------------------------------------------------------------------------------------
def pangenome_reference(args, K, E):
    """The unmodified reference (`index -FD`, `map -ep`) on a scale model of the same construction, and the GPU on
    the same model: -> (cpu_baseline dict, parity dict)."""
    import shutil
    import genmap_b200 as gm
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gmtest as T
    if not T.have_reference():
        raise RuntimeError("oracle/_ref/genmap_ref is not present")
    seqs, stf = gm.synth_pangenome(args.pan_cpu_mbp * 1e6, args.pan_files)
    per = len(seqs) // args.pan_files
    tmp = tempfile.mkdtemp(prefix="gmb_pan_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fdir, odir = os.path.join(tmp, "fa"), os.path.join(tmp, "out")
        os.mkdir(fdir); os.mkdir(odir)
        for g in range(args.pan_files):
            T.write_fasta(os.path.join(fdir, "g%02d.fa" % g), seqs[g * per:(g + 1) * per],
                          names=["g%02d_chr%d" % (g, i + 1) for i in range(per)])
        subprocess.run([T.REF_BIN, "index", "-FD", fdir, "-I", os.path.join(tmp, "index")], check=True, stdout=subprocess.DEVNULL)
        cores = os.cpu_count() or 1
        res = subprocess.run([T.REF_BIN, "map", "-I", os.path.join(tmp, "index"), "-O", odir, "-K", str(K), "-E", str(E), "-ep",
                              "-r", "-fl", "-T", str(cores), "-v"], check=True, stdout=subprocess.PIPE, text=True)
        secs = [float(l.split()[3]) for l in res.stdout.replace("\r", "\n").split("\n") if l.startswith("Mappability computed in")]
        n_all = sum(len(s) for s in seqs)
        lim = np.zeros(len(seqs) + 1, dtype=np.uint64)
        lim[1:] = np.cumsum([len(s) for s in seqs])
        ix = gm.Index.build(seqs, with_sa=True, on_gpu=True, seq_to_file=stf)
        mism, nonuni = 0, 0
        for g in range(args.pan_files):
            s0 = g * per
            tb, tl = int(lim[s0]), int(lim[s0 + per] - lim[s0])
            got = ix.compute_mappability(gm.SearchParams(K, E, True, True, 16), text_begin=tb, text_len=tl,
                                         chrom_cum_lengths=np.ascontiguousarray(lim[s0:s0 + per + 1] - lim[s0]))
            ref = np.fromfile(os.path.join(odir, "g%02d.genmap.freq16" % g), dtype=np.uint16)
            mism += int((got != ref).sum()); nonuni += int((ref > 1).sum())
        ix.close()
        sample = ("`genmap_ref index -FD` + `map -ep` on a scale model of the same construction (%d files x %g Mbp), every "
                  "position of every file, -T %d; time = the sum of its 'Mappability computed in' lines"
                  % (args.pan_files, args.pan_cpu_mbp, cores))
        return ({"value": n_all / sum(secs), "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
                {"positions": int(n_all), "equal": mism == 0, "mismatches": mism, "against": "reference (scale model, all files)",
                 "values_above_1": nonuni})
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main_pangenome(args):
    import torch
    import genmap_b200 as gm
    from genmap_b200 import _lib, parallel
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    K, E = (50, 2) if (args.kmer, args.errors) == (30, 0) else (args.kmer, args.errors)
    workload = "%d x %g Mbp synthetic pan-genome (file g = base + g %% substitutions), K=%d E=%d --exclude-pseudo, both strands, uint16" % (
        args.pan_files, args.pan_file_mbp, K, E)
    if _lib.lib().gmb_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.impl == "reference":
        if rank != 0:
            return 0
        cpu, par = pangenome_reference(args, K, E)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
                          "warmup": 0, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "u32", "data": "synthetic", "config": {"workload": workload, "K": K, "E": E}, "cpu_baseline": cpu,
                          "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return 0
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        bind_to_gpu_numa_node(local)
    nchr = 3
    per_file = int(args.pan_file_mbp * 1e6) // nchr * nchr
    limits = np.arange(args.pan_files * nchr + 1, dtype=np.uint64) * np.uint64(per_file // nchr)
    stf = np.repeat(np.arange(args.pan_files, dtype=np.uint32), nchr)
    t0 = time.time()
    if rank == 0:
        seqs, stf0 = gm.synth_pangenome(args.pan_file_mbp * 1e6, args.pan_files, nchr)
        assert np.array_equal(stf0, stf)
        log("pan-genome generated in %.1f s" % (time.time() - t0))
        t0 = time.time()
        ix = gm.Index.build(seqs, device=local, with_sa=True, on_gpu=True, seq_to_file=stf)
        del seqs
        log("index with suffix array built on GPU in %.1f s, blob %.2f GB" % (time.time() - t0, ix.info.blob_bytes / 1e9))
    bcast_s = None
    if dist is not None:
        blob_t = torch.as_tensor(_DeviceBytes(int(ix.info.device_blob), int(ix.info.blob_bytes)), device=dev) if rank == 0 else None
        dist.barrier()
        t0 = time.time()
        blob_t = parallel.broadcast_blob(blob_t, dist, dev)
        torch.cuda.synchronize()
        bcast_s = time.time() - t0
        log("rank %d: index broadcast over NCCL in %.2f s" % (rank, bcast_s))
        if rank != 0:
            ix = gm.Index.adopt_device(blob_t.data_ptr(), blob_t.numel(), device=local, seq_to_file=stf)
    ix.limits = limits
    ix.set_jump_depth(args.jump_depth)
    stream = torch.cuda.current_stream().cuda_stream
    batch = int((args.batch_mpos or 8.0) * (1 << 20))
    out_dev = torch.zeros(per_file, dtype=torch.int16, device=dev)

    def file_args(g):
        s0 = g * nchr
        return dict(text_begin=int(limits[s0]), text_len=per_file, chrom_cum_lengths=np.ascontiguousarray(limits[s0:s0 + nchr + 1] - limits[s0]))

    # every file's positions are split over the ranks; the steps walk through the files
    sb, se = parallel.shard_range(per_file, rank, world)
    b_ = max(1 << 14, min(batch, se - sb))
    plan = []
    for i in range(args.warmup + args.steps + 1):
        g = i % args.pan_files
        off = sb + ((i // args.pan_files) * b_) % max(1, se - sb - b_ + 1)
        plan.append((g, off, off + b_))
    p_ep, p_plain = gm.SearchParams(K, E, True, True, 16), gm.SearchParams(K, E, True, False, 16)

    def run(g, b, e, p=p_ep, **kw):
        return ix.compute_mappability_device(p, out_dev.data_ptr(), pos_begin=b, pos_end=e, stream=stream, **file_args(g), **kw)

    for g, b, e in plan[:args.warmup]:
        run(g, b, e, sync=False)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    timed = plan[args.warmup:args.warmup + args.steps]
    for g, b, e in timed:
        run(g, b, e, sync=False)
    ev1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    npos = sum(e - b for _, b, e in timed)
    if dist is not None:
        ms = parallel.max_over_ranks(ms, dist, dev)
        npos = parallel.sum_over_ranks(npos, dist, dev)
    value = npos / (ms * 1e-3)
    # end to end: the same batches through the host-buffer call (slice of c into pinned host memory)
    host_t = torch.empty(b_, dtype=torch.int16).pin_memory()
    host = host_t.numpy().view(np.uint16)
    g, b, e = plan[-1]
    ix.compute_mappability_range(p_ep, b, e, out=host, **file_args(g))
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for g, b, e in timed:
        ix.compute_mappability_range(p_ep, b, e, out=host, **file_args(g))
    dt = time.perf_counter() - t0
    if dist is not None:
        dt = parallel.max_over_ranks(dt, dist, dev)
    e2e = {"value": npos / dt, "unit": UNIT, "h2d_bytes_per_step": int(4 * 8 + 4 * len(stf)), "d2h_bytes_per_step": int(2 * b_), "numa": NUMA_NOTE}
    roof = None
    if rank == 0:
        # kernel time per launch; rank-block fetches from the instrumented kernel on the same batches WITHOUT -ep (the same
        # searches; the instrumented instantiation is not built for -ep, whose extra reads — one suffix-array entry per
        # located occurrence — are therefore not in the figure)
        k_ms, f_tot, lut_tot, searched, txt = [], 0, 0, 0, 0
        for g, b, e in timed:
            st = run(g, b, e)
            k_ms.append(st.kernel_ms); searched += int(st.positions)
        p_plain = gm.SearchParams(K, E, True, False, 16, block_kmers=int(st.block_kmers))  # the same blocks as the -ep plan
        for g, b, e in timed:
            st = run(g, b, e, p=p_plain, count_fetches=True)
            f_tot += int(st.rank_block_fetches); lut_tot += int(st.jump_table_reads); txt += int(st.text_reads)
        peak, peak_src = peak_hbm()
        per_launch = f_tot * 32.0 / len(timed)
        achieved = per_launch / (float(np.mean(k_ms)) * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "rank_block_bytes_per_position": f_tot * 32.0 / max(searched, 1),
                "table_reads_per_position": lut_tot / max(searched, 1), "text_reads_per_position": txt / max(searched, 1),
                "kernel_ms_per_launch": float(np.mean(k_ms)), "block_kmers": int(p_plain.block_kmers),
                "note": "fetches counted on the same batches without -ep (same searches); -ep adds one suffix-array read per located occurrence"}
    hbm = {"index_blob_with_sa": int(ix.info.blob_bytes), "jump_tables": int(ix.refresh_info().jump_table_bytes), "result_vector": int(2 * per_file)}
    cpu, par = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            del out_dev
            ix.close()
            torch.cuda.empty_cache()
            cpu, par = pangenome_reference(args, K, E)
            log("reference on the scale model: %.3f M positions/s, parity %s" % (cpu["value"] / 1e6, par))
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (ex,)}
    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
                          "data": "synthetic",
                          "config": {"workload": workload, "K": K, "E": E, "genome_bp": int(limits[-1]), "positions_per_step_per_gpu": b_,
                                     "sharding": "every file's positions range-partitioned over %d GPU(s), index + suffix array replicated (NCCL broadcast%s)"
                                                 % (world, "" if bcast_s is None else ": %.2f s" % bcast_s),
                                     "cache": "inputs larger than L2: %.1f GB index vs 126 MB L2" % (hbm["index_blob_with_sa"] / 1e9),
                                     "hbm_bytes": hbm},
                          "parity": {"pangenome_K%d_E%d_ep" % (K, E): par} if par else None, "clocks": clocks, "e2e": e2e,
                          "gpu_launches": args.steps, "roofline": roof, "cpu_baseline": cpu}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.config == "pangenome":
        return main_pangenome(args)
    import torch
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    K, E = args.kmer, args.errors
    total = int(args.genome_mbp * 1e6)
    workload = "%g Mbp synthetic DNA (%d chr, seed %d), K=%d E=%d, both strands, uint16 counts" % (
        args.genome_mbp, args.nchr, args.seed, K, E)

    if args.impl == "reference" and rank != 0:
        return 0

    import genmap_b200 as gm
    from genmap_b200 import _lib
    if _lib.lib().gmb_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.impl == "ours" and world > 1:
        bind_to_gpu_numa_node(local)
    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic genome + index (rank 0 builds on its GPU, NCCL broadcast to the other ranks) -------
    limits = genome_limits(total, args.nchr)
    n_text = int(limits[-1])
    seqs = None
    t0 = time.time()
    if rank == 0:
        seqs = gm.synth_genome(total, args.nchr, args.seed)
        log("genome generated in %.1f s" % (time.time() - t0))
        t0 = time.time()
        # the suffix array (4 B/base more HBM) is only needed by the CPU arm, to write a reference-format index: the
        # timed index is the one `genmap index --no-sa` builds; the CPU arm builds its own afterwards (index_for_cpu_arm)
        ix = gm.Index.build(seqs, device=local, on_gpu=True, with_sa=args.impl == "reference")
        log("index built on GPU in %.1f s %s, blob %.2f GB" % (time.time() - t0, ix.build_timings_ms, ix.info.blob_bytes / 1e9))
    bcast_s = 0.0
    if dist is not None:
        from genmap_b200 import parallel
        # the library-owned device blob, viewed as a tensor so NCCL can broadcast it in place
        blob_t = torch.as_tensor(_DeviceBytes(int(ix.info.device_blob), int(ix.info.blob_bytes)), device=dev) if rank == 0 else None
        dist.barrier()  # (the other ranks have been waiting for rank 0's genome and index: not part of the broadcast)
        torch.cuda.synchronize()
        t0 = time.time()
        blob_t = parallel.broadcast_blob(blob_t, dist, dev)
        torch.cuda.synchronize()
        bcast_s = time.time() - t0
        log("rank %d: index broadcast over NCCL in %.2f s" % (rank, bcast_s))
        if rank != 0:
            ix = gm.Index.adopt_device(blob_t.data_ptr(), blob_t.numel(), device=local)
    ix.limits = limits
    ix.set_jump_depth(args.jump_depth)

    params = gm.SearchParams(K, E)
    stream = torch.cuda.current_stream().cuda_stream

    if args.impl == "reference":
        arm = make_cpu_arm(seqs, ix)
        ix.close()
        torch.cuda.empty_cache()
        rate, times, npos = cpu_rate(arm, K, E, args.ref_step_seconds, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": workload, "K": K, "E": E, "genome_bp": n_text, "positions_per_step": npos},
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.sample(npos, K)},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        arm.close()
        print(json.dumps(line), flush=True)
        return 0

    # ---- one measurement = (E, batch) ---------------------------------------------------------------------
    from genmap_b200 import parallel
    shard_b, shard_e = parallel.shard_range(n_text, rank, world)
    out_dev = torch.zeros(n_text, dtype=torch.int16, device=dev)

    def batches(E_, batch, count):
        return parallel.step_batches(shard_b, shard_e, batch, count)

    def measure(E_, batch, steps, warmup, with_e2e, sample_clocks):
        p = gm.SearchParams(K, E_)
        bl = batches(E_, batch, warmup + steps)
        for b, e in bl[:warmup]:
            ix.compute_mappability_device(p, out_dev.data_ptr(), pos_begin=b, pos_end=e, stream=stream, sync=False)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for b, e in bl[warmup:]:
            ix.compute_mappability_device(p, out_dev.data_ptr(), pos_begin=b, pos_end=e, stream=stream, sync=False)
        ev1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = ev0.elapsed_time(ev1)
        npos = sum(e - b for b, e in bl[warmup:])
        if dist is not None:
            ms = parallel.max_over_ranks(ms, dist, dev)
            npos_all = parallel.sum_over_ranks(npos, dist, dev)
        else:
            npos_all = npos
        res = {"ms": ms, "positions": npos_all, "value": npos_all / (ms * 1e-3), "clocks": clocks, "batches": bl[warmup:]}
        # rank-block fetches of exactly these batches (instrumented kernel, outside the timed region) and
        # per-launch kernel time from CUDA events on the launching stream
        if rank == 0:
            f_tot, k_ms, pos0, lut_tot, jd = 0, [], 0, 0, 0
            for b, e in bl[warmup:]:
                st = ix.compute_mappability_device(p, out_dev.data_ptr(), pos_begin=b, pos_end=e, stream=stream, count_fetches=True)
                f_tot += int(st.rank_block_fetches); lut_tot += int(st.jump_table_reads); jd = int(st.jump_depth)
            for b, e in bl[warmup:]:
                st = ix.compute_mappability_device(p, out_dev.data_ptr(), pos_begin=b, pos_end=e, stream=stream)
                k_ms.append(st.kernel_ms); pos0 += int(st.positions)
            res.update(fetches=f_tot, kernel_ms=float(np.mean(k_ms)), searched=pos0, lut_reads=lut_tot, jump_depth=jd)
        if with_e2e:
            host_t = torch.empty(batch, dtype=torch.int16).pin_memory()
            host = host_t.numpy().view(np.uint16)
            bl2 = batches(E_, batch, 1 + steps)
            ix.compute_mappability_range(p, bl2[0][0], bl2[0][1], out=host)
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            for b, e in bl2[1:]:
                ix.compute_mappability_range(p, b, e, out=host)
            dt = time.perf_counter() - t0
            npos2 = sum(e - b for b, e in bl2[1:])
            if dist is not None:
                dt = parallel.max_over_ranks(dt, dist, dev)
                npos2 = parallel.sum_over_ranks(npos2, dist, dev)
            # per step: work-range table + step table + counters go H2D, the slice of c comes back D2H
            # what the host link alone allows: the same bytes device -> pinned host with no search at all, every rank at
            # once (the e2e figure cannot exceed it; at E = 0 it is what bounds it)
            src = out_dev[:batch]
            host_t.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                host_t.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            dt_copy = time.perf_counter() - t0
            if dist is not None:
                dt_copy = parallel.max_over_ranks(dt_copy, dist, dev)
            # per step: the work-range table goes H2D (search tables are cached in the handle), the slice of c comes back
            res["e2e"] = {"value": npos2 / dt, "unit": UNIT,
                          "h2d_bytes_per_step": int(3 * 8 + 8),
                          "d2h_bytes_per_step": int(2 * batch),
                          "d2h_only_gb_per_s_per_gpu": 2.0 * batch * steps / dt_copy / 1e9,
                          "d2h_only_positions_per_s": world * batch * steps / dt_copy,
                          "numa": NUMA_NOTE}
        return res

    batch = int((args.batch_mpos or default_batch(E)) * (1 << 20))
    batch = max(1 << 16, min(batch, shard_e - shard_b))
    main_r = measure(E, batch, args.steps, args.warmup, True, True)
    peak, peak_src = peak_hbm()

    def roofline(r):
        if "fetches" not in r:
            return None
        blk = float(ix.info.rank_block_bytes)  # 32: one sector per rank boundary
        per_launch = r["fetches"] * blk / len(r["batches"])
        achieved = per_launch / (r["kernel_ms"] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f)
            key = "K%d_E%d_batch%d_genome%d" % (K, r["E"], r["batch"], n_text)
            traffic = tr.get(key)
            traffic_src = tr.get("_source")
            if isinstance(traffic, dict):  # {"bytes": per-launch DRAM bytes, "git": sha of the kernel measured, "csv": file}
                traffic_src = "%s @ %s" % (traffic.get("csv"), traffic.get("git"))
                traffic = traffic.get("bytes")
        except Exception:
            traffic_src = None
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "rank_block_bytes": blk,
                "rank_block_bytes_per_position": r["fetches"] * blk / max(r["searched"], 1),
                # stricter figure: + jump-table entries (16 B) + pattern text (K/4 -> 16 B) + result (2 B)
                "total_algorithmic_bytes_per_position": (r["fetches"] * blk + r["lut_reads"] * 16.0) / max(r["searched"], 1) + 18.0,
                # one dependent random request per rank block / jump-table entry; for scale, the best row of the
                # pointer-chase microbenchmark (tools/randread.cu, profiles/r01/s1_randread.txt: 2048 threads/SM x 4 loads
                # in flight) reaches 45.8 G 32-byte hops/s — a reference point, not a ceiling of this kernel
                "random_requests_per_s": (r["fetches"] + r["lut_reads"]) / len(r["batches"]) / (r["kernel_ms"] * 1e-3),
                "random_request_rate_best_microbenchmark_per_s": 45.8e9,
                # what the DRAM actually does: 64-byte accesses per second (ncu traffic / 64 B / kernel time)
                "dram_64B_accesses_per_s": (traffic / 64.0 / (r["kernel_ms"] * 1e-3)) if traffic else None,
                # ... and as a fraction of the copy peak: what the memory system is really asked to move (the north-star
                # fraction above counts 32 useful bytes per rank block only, and falls when a change removes fetches)
                "dram_gbs": (traffic / (r["kernel_ms"] * 1e-3) / 1e9) if traffic else None,
                "dram_frac": (traffic / (r["kernel_ms"] * 1e-3) / 1e9 / peak) if traffic else None,
                # the same fraction on the stricter figure (rank blocks + 16-byte jump-table entries)
                "frac_with_table_reads": (r["fetches"] * blk + r["lut_reads"] * 16.0) / len(r["batches"]) / (r["kernel_ms"] * 1e-3) / 1e9 / peak,
                "jump_table_depth": r["jump_depth"], "kernel_ms_per_launch": r["kernel_ms"]}

    main_r["E"], main_r["batch"] = E, batch
    extras = {}
    if rank == 0 or dist is not None:
        for e_s in [x for x in args.extras.split(",") if x.strip() != ""]:
            E2 = int(e_s)
            if E2 == E:
                continue
            b2 = int(default_batch(E2) * (1 << 20))
            b2 = max(1 << 14, min(b2, shard_e - shard_b))
            try:
                r2 = measure(E2, b2, max(3, args.steps // 2), 3, True, False)
            except Exception as ex:  # an extra must never cost the headline line (single process only: with several
                if dist is not None:  # ranks a one-sided failure would leave the others waiting in a collective)
                    raise
                extras["K%d_E%d" % (K, E2)] = {"value": None, "unit": UNIT, "error": repr(ex)}
                continue
            r2["E"], r2["batch"] = E2, b2
            extras["K%d_E%d" % (K, E2)] = {"value": r2["value"], "unit": UNIT, "ms_per_step": r2["ms"] / max(3, args.steps // 2),
                                          "positions_per_step": b2, "e2e": r2.get("e2e"), "roofline": roofline(r2)}

    cpu = None
    parity = {}
    index_gb = ix.info.blob_bytes / 1e9
    hbm = {"index_blob": int(ix.info.blob_bytes), "jump_tables": int(ix.refresh_info().jump_table_bytes),
           "result_vector": int(2 * n_text)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            del out_dev
            ix = index_for_cpu_arm(seqs, ix, local)
            arm = make_cpu_arm(seqs, ix)
            rate, _, npos_cpu = cpu_rate(arm, K, E, args.cpu_seconds)
            cpu = {"value": rate, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.sample(npos_cpu, K)}
            parity["K%d_E%d" % (K, E)] = parity_check(arm, ix, K, E)
            log("parity %s" % (parity,))
            for name, ex in extras.items():  # the same CPU arm for the other (K,E) lines, shorter samples
                if ex.get("value") is None:
                    continue
                E2 = int(name.split("_E")[1])
                r2, _, n2 = cpu_rate(arm, K, E2, args.cpu_seconds / 2)
                ex["cpu_baseline"] = {"value": r2, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.sample(n2, K)}
                parity[name] = parity_check(arm, ix, K, E2)
                log("parity %s: %s" % (name, parity[name]))
            arm.close()
        except Exception as ex:  # the baseline is a reported extra; never lose the GPU line over it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        line = {"metric": METRIC, "value": main_r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_r["ms"] / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": workload, "K": K, "E": E, "genome_bp": n_text, "positions_per_step_per_gpu": batch,
                           "sharding": "positions range-partitioned over %d GPU(s), index replicated (NCCL broadcast%s)"
                                       % (world, ": %.2f s" % bcast_s if world > 1 else ""),
                           "cache": "inputs larger than L2: %.2f GB index vs 126 MB L2, every step searches different positions"
                                    % index_gb,
                           "hbm_bytes": hbm},
                "parity": parity or None,
                "clocks": main_r["clocks"], "e2e": main_r.get("e2e"), "gpu_launches": args.steps,
                "roofline": roofline(main_r), "cpu_baseline": cpu, "extra": extras}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
