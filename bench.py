#!/usr/bin/env python
"""bench.py -- stLFR read pairs classified per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

The JSON line's top level is the contract's line on BASELINE.json configs[1] (cfg2): a synthetic 100 Mbp diploid
trio, k=21 parent-unique k-mers, 20 M stLFR 100 bp read pairs over 500 k barcodes; one *step* = one pass of the hot
path (pack -> canonical k-mers -> table probe -> per-barcode reduce, then the count collection) over the rank's
whole workload.  At N>1 every rank classifies its own 20 M pairs of the same trio (weak scaling), the k-mer table
is replicated per GPU and the per-barcode partial counts are summed with ONE ncclReduce inside hast_finish.

value        device-resident inputs, CUDA-event time on the launching stream, max over ranks
e2e          the same step through hast_submit_batch with HOST (pinned) buffers, H2D copies and the D2H read of
             the counts inside the timed region
roofline     fused kernel: algorithmic bytes (32 B = one table sector per k-mer lookup) / kernel time vs the
             measured HBM peak, next to the DRAM bytes ncu measured for the same launch and the random-gather rate
cpu_baseline the UNTOUCHED reference binary (oracle/_ref/classify_O2, all host threads) on a bounded sample of the
             same workload, k-mer load time subtracted
parity       (in the run that is timed) a barcode-complete subsample of the timed counts against the CPU oracle

Two more legs ride on the same line:

cfg3         BASELINE.json configs[2] at its stated shape as a STRONG-scaling job: 3.1 Gbp trio (~62 M parent-unique
             21-mers, 1 GiB table), 600 M read pairs / N per rank generated on the device from (seed, pair index),
             20 M barcodes (160 MB of counters per GPU), the ncclReduce and the read-back broken out, and an in-run
             N-GPU parity check (sum of the per-rank partials == sum of the reduced array; a barcode-complete
             subsample of the REDUCED counts == the CPU oracle on those reads)
cli          (N=1) the drop-in process: bin/classify against the untouched reference binary on the same plain and
             gzip FASTQ files, whole-process wall and streaming rate, tables compared byte for byte
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "stLFR read pairs classified/s"
UNIT = "pairs/s"
ALG_BYTES_PER_LOOKUP = 32            # one sector-aligned 4-slot bucket
SUB_BATCH_READS = 4_000_000          # reads per hast_submit_batch (< 4 GiB of bases each)


# stdout carries exactly ONE JSON line.  Libraries underneath (NCCL prints its version to stdout when
# NCCL_DEBUG is set on the box) write to file descriptor 1, so keep a private copy of the real stdout
# for the result line and point fd 1 at stderr for everything else.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def workload_spec(name: str, rank: int):
    from hast_b200 import synth
    spec = synth.config(name)
    spec.read_seed = 51 + rank
    return spec


WORKLOAD_TEXT = {
    "cfg2": "configs[1]: synthetic 100 Mbp diploid trio (0.1% het), k=21, 20M stLFR 100bp read pairs / 500k barcodes",
    "cfg3t": "configs[2] table scale: human-size parent-unique k-mer lists (62 M keys: the cfg2 trio + random decoys), "
             "k=21, 20M stLFR 100bp read pairs / 500k barcodes per GPU",
    "cfg3": "configs[2], the shard of one GPU out of eight: human-size parent-unique k-mer lists (62 M keys: the cfg2 trio + "
            "random decoys), k=21, 75M stLFR 100bp read pairs (600M / 8) over 20M barcodes",
    "cfg1": "configs[0]: synthetic 5 Mbp diploid trio (0.1% het), k=21, 200k stLFR 100bp read pairs / 10k barcodes",
    "small": "dev: 500 kbp trio, 20k pairs / 1k barcodes",
}


# ----------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ----------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs: NVML in-process every 2 ms
    (the timed region is only tens of milliseconds), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.source = "nvidia-smi"
        self._stop = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None
        self._t = threading.Thread(target=self._run_nvml if self._nvml else self._run_smi, daemon=True)

    def _run_nvml(self):
        nv = self._nvml
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for n, b in bits.items():
                    if r & b:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.002)

    def _run_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "source": self.source}


# ----------------------------------------------------------------------------
# the reference's CPU implementation on a bounded sample
# ----------------------------------------------------------------------------
ADAPTOR_F = b"CTGTCTCTTATACACATCTTAGGAAGACAAGCACTGACGACATGA"     # classify.cpp:312
ADAPTOR_R = b"TCTGCTGAGTCGAGAACGTCTCTGTGAGCCAAGGAGTTGCTCTGG"     # classify.cpp:313


class RefRunner:
    """The UNTOUCHED reference binary (oracle/_ref/classify_O2, built from /root/reference sources by
    oracle/Makefile) on the first sample_pairs pairs of the workload, written as FASTQ.  pairs/s is net of
    the k-mer load time, measured once with a one-read FASTQ.  Falls back to the
    plain-C oracle port when the reference binary is absent."""

    def __init__(self, trio, sample_pairs: int, workdir: Path, gz_too: bool = False):
        self.trio, self.sample = trio, min(sample_pairs, trio.spec.n_pairs)
        self.ref = ROOT / "oracle" / "_ref" / "classify_O2"
        self.cores = os.cpu_count() or 1
        self.t_load = None
        self.table = None                    # stdout of the last reference run on the plain files
        self.gz = None
        if self.ref.exists():
            self.pat, self.mat = trio.write_kmer_lists(workdir)
            self.r1, self.r2 = trio.write_fastq(workdir, gz=False, lo=0, hi=self.sample, stem="sample")
            self.one, _ = trio.write_fastq(workdir, gz=False, lo=0, hi=1, stem="one")
            if gz_too:                       # one gzip member per file, level 6: what a sequencer ships
                self.gz = trio.write_fastq(workdir / "gz", gz=6, lo=0, hi=self.sample, stem="sample")
            os.sync()                        # timed legs read clean page-cache pages, not files still being written back
        else:
            sys.path.insert(0, str(ROOT / "tests"))
            import oracle as orc
            self.o = orc.Oracle()
            self.o.load_kmers_packed(trio.pat, trio.spec.k, 0)
            self.o.load_kmers_packed(trio.mat, trio.spec.k, 1)
            self.o.init_adaptor()
            self.batch = trio.batch(0, self.sample)
        self.threads = None

    def cmd(self, threads, reads):
        c = [str(self.ref), "--hap0", self.pat, "--hap1", self.mat, "--weight0", "1.04", "--thread", str(threads)]
        for r in reads:
            c += ["--read", r]
        return c

    def timed(self, cmd):
        t = time.perf_counter()
        r = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        return time.perf_counter() - t, r.stdout

    def run(self):
        if not self.ref.exists():
            bases, off, bc = self.batch
            t = time.perf_counter()
            self.o.classify_batch(bases, off.astype(np.uint64), bc, self.trio.n_barcodes, nthreads=self.cores)
            net = time.perf_counter() - t
            return {"value": self.sample / net, "unit": UNIT, "cores": self.cores, "kind": "port",
                    "sample": f"first {self.sample} pairs in memory; oracle/hast_oracle.c ho_classify_batch on {self.cores} threads"}
        if self.t_load is None:
            self.t_load, _ = self.timed(self.cmd(8, [self.one]))
        if self.threads is None:             # the reference stops scaling early (one reader thread,
            best = None                      # classify.cpp:257-269): sweep and keep the best, BASELINE.md 3.2
            for t in sorted({min(8, self.cores), min(16, self.cores), min(32, self.cores), self.cores}):
                dt, out = self.timed(self.cmd(t, [self.r1, self.r2]))
                log(f"reference --thread {t}: {dt:.2f}s")
                if best is None or dt < best[1]:
                    best = (t, dt, out)
            self.threads, t_all, self.table = best
        else:
            t_all, self.table = self.timed(self.cmd(self.threads, [self.r1, self.r2]))
        net = max(t_all - self.t_load, 1e-6)
        return {"value": self.sample / net, "unit": UNIT, "cores": self.threads, "host_cores": self.cores,
                "kind": "reference",
                "sample": f"first {self.sample} pairs as plain FASTQ; oracle/_ref/classify_O2 with {self.threads} threads "
                          f"(best of sweep); wall {t_all:.2f}s minus k-mer load {self.t_load:.2f}s",
                "wall_s": t_all, "kmer_load_s": self.t_load}


def reference_arm(args):
    """--impl reference: the reference's own CPU path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from hast_b200 import synth
    cores = os.cpu_count() or 1
    spec = workload_spec(args.workload, 0)
    # the reference arm needs the k-mer lists of the full trio but only a sample of the reads
    sample = min(spec.n_pairs, max(100_000, min(4_000_000, 125_000 * cores)))
    spec.n_pairs = sample
    log(f"reference arm: generating trio ({args.workload}, {sample} pairs sample) ...")
    dev = "cuda" if _torch_cuda() else "cpu"
    trio = synth.make_trio(spec, device=dev)
    vals = []
    with tempfile.TemporaryDirectory(prefix="hast_ref_", dir=os.environ.get("TMPDIR", "/tmp")) as d:
        runner = RefRunner(trio, sample, Path(d))
        res = None
        for i in range(args.warmup + args.steps):
            res = runner.run()
            if i >= args.warmup:
                vals.append(res["value"])
            log(f"reference step {i}: {res['value']:.0f} pairs/s")
    v = float(np.mean(vals))
    res["value"] = v
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload], "k": spec.k, "read_len": spec.read_len,
                       "sample_pairs": sample},
            "cpu_baseline": res,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def _torch_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ----------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------
class Runtime:
    """rank / device / collectives of this process (one process per GPU, torchrun env)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the classification path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = f"cuda:{self.local}"
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.dev))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        t = self.torch.tensor(list(xs), dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(v) for v in t.tolist()]

    def comm_for(self, eng):
        if self.world > 1:
            uid = [eng.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(uid, src=0)
            eng.comm_init_rank(self.world, self.rank, uid[0])

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


class _DevView:
    """zero-copy torch view of a raw device pointer (the engine's int32 counts[n][2])"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n, 2), "typestr": "<i4", "data": (ptr, False), "version": 2}


def local_count_sums(R, eng, nb):
    """column sums of THIS rank's partial counts, read on the device (before / independent of the reduce)"""
    ptr, n = eng.counts_device_ptr()
    eng.sync()
    v = R.torch.as_tensor(_DevView(ptr, min(n, nb)), device=R.dev)
    s = v.to(R.torch.int64).sum(0)
    return int(s[0].item()), int(s[1].item())


def load_peaks():
    pk = ROOT / "MEASURED_PEAKS.json"
    peaks = json.loads(pk.read_text()) if pk.exists() else {}
    return float(peaks.get("hbm_gbs", 6650.0)), ("MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s")


def measured_traffic(key):
    """DRAM bytes per launch of the fused kernel from the round's `ncu --set full` capture (profiles/traffic.json)"""
    tf = ROOT / "profiles" / "traffic.json"
    try:
        d = json.loads(tf.read_text())
        return d.get(key), d.get("capture")
    except Exception:
        return None, None


def oracle_for(pat, mat, k):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle as orc
    o = orc.Oracle()
    o.load_kmers_packed(pat, k, 0)
    o.load_kmers_packed(mat, k, 1)
    o.init_adaptor(ADAPTOR_F, ADAPTOR_R)
    return o


def build_table(eng, k, pat, mat, scale=1.0):
    n_keys = pat.size + mat.size
    eng.table_begin(k, int(n_keys * scale))
    t0 = time.perf_counter()
    eng.table_add_packed(pat, 0)
    eng.table_add_packed(mat, 1)
    eng.table_erase_seq(ADAPTOR_F)
    eng.table_erase_seq(ADAPTOR_R)
    info = eng.table_info()
    return info, time.perf_counter() - t0


def roofline_of(lookups_per_launch, ms_per_launch, peak, peak_source, traffic_key, gather_4g, gather_tbl, info):
    achieved = lookups_per_launch * ALG_BYTES_PER_LOOKUP / (ms_per_launch * 1e-3) / 1e9
    traffic, capture = measured_traffic(traffic_key)
    out = {"bound": "alu-issue (L2-resident pre-filter answers ~95 % of the lookups; only the exact-table probes reach HBM)",
           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "frac_algorithmic": achieved / peak,
           "traffic": traffic, "traffic_capture": capture,
           "frac_dram": (traffic / (ms_per_launch * 1e-3) / 1e9 / peak) if traffic else None,
           "ms_per_launch": ms_per_launch, "lookups_per_launch": int(lookups_per_launch),
           "lookups_per_s": lookups_per_launch / (ms_per_launch * 1e-3),
           "peak_source": peak_source,
           "random_gather_gbs_4GiB": gather_4g, "random_gather_gbs_table_span": gather_tbl,
           "frac_of_random_gather": (achieved / gather_4g) if gather_4g else None,
           "table_bytes": int(info.bytes), "filter_bytes": int(info.filter_bytes),
           "note": "achieved/frac are ALGORITHMIC (32 B per lookup, SURVEY.md 8d): the kernel does not move those bytes "
                   "-- frac_dram is what it really pulls from HBM (ncu dram__bytes per launch / CUDA-event time / peak)"}
    return out


def leg_cfg2(args, R):
    """The contract's line: configs[1] per rank, device-resident `value`, `e2e`, roofline, in-run parity."""
    torch = R.torch
    from hast_b200 import synth
    from hast_b200.capi import Engine
    rank, world, dev = R.rank, R.world, R.dev
    spec = workload_spec(args.workload, rank)
    t0 = time.perf_counter()
    trio = synth.make_trio(spec, device=dev, keep_reads_on_device=True)
    torch.cuda.synchronize()
    log(f"rank {rank}: trio generated in {time.perf_counter() - t0:.1f}s: pat {trio.pat.size} mat {trio.mat.size} "
        f"k-mers, {spec.n_pairs} pairs, {trio.n_barcodes} barcodes")
    L, P = spec.read_len, spec.n_pairs
    n_reads = 2 * P
    # r1 and r2 are the two halves of one contiguous [2P, L] tensor
    assert trio.r2.data_ptr() == trio.r1.data_ptr() + P * L
    d_bases = torch.as_strided(trio.r1, (n_reads * L,), (1,))
    bc_np = np.concatenate([trio.pair_bc, trio.pair_bc]).astype(np.int32)
    d_bc = torch.from_numpy(bc_np).to(dev)
    sub = min(SUB_BATCH_READS, n_reads)
    d_off = (torch.arange(sub + 1, dtype=torch.int64, device=dev) * L).to(torch.int32)
    torch.cuda.synchronize()
    nb = trio.n_barcodes

    eng = Engine(R.local)
    eng.set_option("kernel", args.kernel)
    eng.set_option("filter_bits_per_key", args.filter_bits)
    eng.set_option("filter_max_bytes", args.filter_max_mib << 20)
    if args.l2_fetch:
        eng.set_option("l2_fetch_granularity", args.l2_fetch)
    if args.reads_per_tile:
        eng.set_option("reads_per_tile", args.reads_per_tile)
    if args.l2_persist_mib >= 0:
        eng.set_option("l2_persist_bytes", args.l2_persist_mib << 20)
    info, t_table = build_table(eng, spec.k, trio.pat, trio.mat, args.table_scale)
    log(f"rank {rank}: table {info.bytes / 2**20:.0f} MiB + pre-filter {info.filter_bytes / 2**20:.1f} MiB, {info.n_entries} entries, "
        f"{info.n_overflow_buckets} overflow buckets, {info.n_displaced} displaced, built in {t_table:.2f}s")
    eng.reserve_barcodes(nb)
    R.comm_for(eng)

    batches = []
    for lo in range(0, n_reads, sub):
        n = min(sub, n_reads - lo)
        batches.append((d_bases.data_ptr() + lo * L, n * L, d_off.data_ptr(), d_bc.data_ptr() + 4 * lo, n))

    def step_device(collect=True):
        for b in batches:
            eng.submit_batch_device(*b)
        if collect:
            return eng.finish(nb, want_counts=(rank == 0))

    # ---- device-resident: `value` ----------------------------------------------
    for _ in range(args.warmup):
        eng.reset_counts()
        step_device()
    eng.reset_counts()
    launches0 = eng.stats()["kernel_launches"]
    R.barrier()
    with ClockSampler(R.local) as clk:
        eng.timer_start()
        for _ in range(args.steps):
            counts = step_device()
        ms_total = eng.timer_stop()
        R.barrier()
    st = eng.stats()
    gpu_launches = st["kernel_launches"] - launches0
    ms_step = R.max_over_ranks(ms_total / args.steps)
    lookups_step = st["lookups"] // args.steps
    value = world * P / (ms_step * 1e-3)

    # ---- parity of the run that was just timed -----------------------------------
    # counts = `steps` identical passes accumulated.  (a) every rank's partial sums add up to the reduced
    # array's sums; (b) N=1: every read of 64 barcodes through the CPU oracle == counts / steps.
    loc = local_count_sums(R, eng, nb)
    tot = R.sum_over_ranks(loc)
    parity = {"checked": False}
    if rank == 0:
        red = counts.astype(np.int64).sum(0)
        parity = {"checked": True, "sum_of_partials_equals_reduced": [int(red[0]), int(red[1])] == tot,
                  "reduced_sums": [int(red[0]), int(red[1])]}
        ok = parity["sum_of_partials_equals_reduced"]
        if world == 1:
            t0 = time.perf_counter()
            rng = np.random.default_rng(7)
            ids = np.sort(rng.choice(nb, size=min(64, nb), replace=False))
            sel = np.nonzero(np.isin(trio.pair_bc, ids))[0]
            rows = torch.from_numpy(np.concatenate([sel, sel + P])).to(dev)
            sb = trio.r1.new_empty((rows.numel(), L))
            torch.index_select(torch.as_strided(trio.r1, (n_reads, L), (L, 1)), 0, rows, out=sb)
            o = oracle_for(trio.pat, trio.mat, spec.k)
            s_off = np.arange(rows.numel() + 1, dtype=np.uint64) * np.uint64(L)
            s_bc = np.concatenate([trio.pair_bc[sel], trio.pair_bc[sel]]).astype(np.uint32)
            want, _ = o.classify_batch(sb.cpu().numpy().reshape(-1), s_off, s_bc, nb, nthreads=os.cpu_count() or 4)
            same = bool((counts[ids].astype(np.int64) == args.steps * want[ids].astype(np.int64)).all())
            sizes_ok = (int(info.size[0]), int(info.size[1])) == (o.set_size(0), o.set_size(1))
            parity.update({"oracle": "oracle/hast_oracle.c ho_classify_batch, full k-mer lists",
                           "oracle_barcodes": int(ids.size), "oracle_reads": int(rows.numel()),
                           "oracle_hits": int(want[ids].sum()), "oracle_counts_identical": same,
                           "oracle_set_sizes_identical": sizes_ok, "seconds": time.perf_counter() - t0})
            ok = ok and same and sizes_ok and want[ids].sum() > 0
            o.close()
        parity["ok"] = bool(ok)
        if not ok:
            raise SystemExit(f"PARITY FAILURE in the timed run: {parity}")

    # kernel-only duration for the roofline (no finish / reduce / D2H in the region)
    eng.reset_counts()
    R.barrier()
    eng.timer_start()
    for _ in range(args.steps):
        step_device(collect=False)
    ms_kernel = eng.timer_stop() / args.steps
    eng.sync()
    peak, peak_source = load_peaks()
    gather = eng.gather_roofline(1 << 28, 4 << 30) if rank == 0 else None
    gather_tbl = eng.gather_roofline(1 << 28, info.bytes) if rank == 0 else None
    roofline = roofline_of(lookups_step / len(batches), ms_kernel / len(batches), peak, peak_source,
                           "fused_kernel_dram_bytes_per_launch", gather, gather_tbl, info)
    roofline.update({"kernel": "classify_kernel" if args.kernel >= 1 else "tile_kernel<MODE_CLASSIFY>",
                     "launches_per_step": len(batches), "kernel_ms_per_step": ms_kernel,
                     "kernel_share_of_step": ms_kernel / (ms_total / args.steps)})

    # ---- end to end through the host-buffer ABI: `e2e` -------------------------
    e2e = None
    if not args.no_e2e:
        h_bases = torch.empty(n_reads * L, dtype=torch.uint8, pin_memory=True)
        h_bases.copy_(d_bases)
        h_bc = torch.from_numpy(bc_np).pin_memory()
        h_off = d_off.cpu().pin_memory()
        torch.cuda.synchronize()
        hb = [(h_bases.data_ptr() + lo * L, min(sub, n_reads - lo) * L, h_off.data_ptr(), h_bc.data_ptr() + 4 * lo,
               min(sub, n_reads - lo)) for lo in range(0, n_reads, sub)]

        def step_host():
            for b in hb:
                eng.submit_batch_ptr(*b)
            return eng.finish(nb, want_counts=(rank == 0))

        for _ in range(2):
            eng.reset_counts()
            c_h = step_host()
        eng.reset_counts()
        R.barrier()
        t = time.perf_counter()
        for _ in range(args.steps):
            c_h = step_host()
        eng.sync()
        dt = R.max_over_ranks((time.perf_counter() - t) / args.steps)
        R.barrier()
        if rank == 0:                      # both legs accumulated `steps` identical passes
            assert (c_h == counts).all(), "host-buffer path and device-resident path disagree"
        h2d = n_reads * L + sum(b[4] + 1 for b in hb) * 4 + n_reads * 4
        e2e = {"value": world * P / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(nb * 8), "ms_per_step": dt * 1e3,
               "h2d_gbs_per_gpu": h2d / dt / 1e9,        # against ~55 GB/s of a PCIe Gen5 x16 link: this leg is PCIe-bound
               "path": "hast_submit_batch (pinned host buffers, double-buffered cudaMemcpyAsync) + hast_finish",
               "counts_identical_to_device_path": True}

        # ---- the same ASCII host buffers, packed to 2 bits by the library on the host's cores (option host_pack_threads):
        #      pack + H2D of a quarter of the bytes + kernel + read-back, all inside the timed region ----
        if args.kernel >= 1 and args.host_pack_threads != 0:
            cores = os.cpu_count() or 1
            hp = args.host_pack_threads if args.host_pack_threads > 0 else max(1, cores // world - 2)
            eng.set_option("host_pack_threads", hp)

            def step_hp():
                for b in hb:
                    eng.submit_batch_ptr(*b)
                return eng.finish(nb, want_counts=(rank == 0))

            for _ in range(2):
                eng.reset_counts()
                c_q = step_hp()
            eng.reset_counts()
            R.barrier()
            t = time.perf_counter()
            for _ in range(args.steps):
                c_q = step_hp()
            eng.sync()
            dtq = R.max_over_ranks((time.perf_counter() - t) / args.steps)
            R.barrier()
            eng.set_option("host_pack_threads", 0)
            if rank == 0:
                assert (c_q == counts).all(), "host-packed path and device-resident path disagree"
            h2d_q = (n_reads * L + 15) // 16 * 4 + sum(b[4] + 1 for b in hb) * 4 + n_reads * 4 + (n_reads + 31) // 32 * 4
            e2e["host_packed"] = {"value": world * P / dtq, "unit": UNIT, "h2d_bytes_per_step": int(h2d_q),
                                  "d2h_bytes_per_step": int(nb * 8), "ms_per_step": dtq * 1e3, "host_pack_threads": hp,
                                  "path": "hast_submit_batch with option host_pack_threads: the same ASCII host buffers, packed to "
                                          "2 bits on the host's cores inside the call (and the timed region), a quarter of the bytes copied"}
        del h_bases

        # ---- the same step with batches as the C++ parser hands them over by default: 2-bit packed ----
        if args.kernel >= 1:
            shifts = (30 - 2 * torch.arange(16, device=dev, dtype=torch.int64))
            pk_w, pk_f = [], []
            for lo in range(0, n_reads, sub):
                n = min(sub, n_reads - lo)
                seg = d_bases[lo * L:(lo + n) * L]
                pad = (-seg.numel()) % 16
                codes = ((seg >> 1) & 3)
                if pad:
                    codes = torch.cat([codes, torch.zeros(pad, dtype=torch.uint8, device=dev)])
                w = torch.empty(codes.numel() // 16, dtype=torch.int32, device=dev)
                CH = 1 << 24
                c16 = codes.view(-1, 16)
                for a in range(0, c16.shape[0], CH):
                    w[a:a + CH] = ((c16[a:a + CH].to(torch.int64) << shifts).sum(1) & 0xFFFFFFFF).to(torch.int32)
                flag = (seg.view(n, L) == ord("N")).any(1)
                fpad = (-n) % 32
                if fpad:
                    flag = torch.cat([flag, torch.zeros(fpad, dtype=torch.bool, device=dev)])
                fw = ((flag.view(-1, 32).to(torch.int64) << torch.arange(32, device=dev, dtype=torch.int64)).sum(1)
                      & 0xFFFFFFFF).to(torch.int32)
                pk_w.append(w.cpu().pin_memory())
                pk_f.append(fw.cpu().pin_memory())
                del codes, c16, w, flag, fw
            torch.cuda.synchronize()
            pb = [(pk_w[i].data_ptr(), min(sub, n_reads - lo) * L, h_off.data_ptr(), h_bc.data_ptr() + 4 * lo,
                   pk_f[i].data_ptr(), min(sub, n_reads - lo)) for i, lo in enumerate(range(0, n_reads, sub))]

            def step_packed():
                for b in pb:
                    eng.submit_batch_packed_ptr(*b)
                return eng.finish(nb, want_counts=(rank == 0))

            for _ in range(2):
                eng.reset_counts()
                c_p = step_packed()
            eng.reset_counts()
            R.barrier()
            t = time.perf_counter()
            for _ in range(args.steps):
                c_p = step_packed()
            eng.sync()
            dtp = R.max_over_ranks((time.perf_counter() - t) / args.steps)
            R.barrier()
            if rank == 0:
                assert (c_p == counts).all(), "packed host-buffer path and device-resident path disagree"
            h2d_p = sum(w.numel() * 4 for w in pk_w) + sum(f.numel() * 4 for f in pk_f) + sum(b[5] + 1 for b in pb) * 4 \
                + n_reads * 4
            e2e["packed"] = {"value": world * P / dtp, "unit": UNIT, "h2d_bytes_per_step": int(h2d_p),
                             "d2h_bytes_per_step": int(nb * 8), "ms_per_step": dtp * 1e3,
                             "h2d_gbs_per_gpu": h2d_p / dtp / 1e9,
                             "path": "hast_submit_batch_packed: 2-bit words + containN bits as bin/classify's parser "
                                     "emits them by default (packing happens while parsing, outside this region)"}
            del pk_w, pk_f
        ceil = h2d_ceiling(world)
        if ceil:
            e2e["h2d_ceiling_gbs_per_gpu"] = ceil
            e2e["frac_of_h2d_ceiling"] = e2e["h2d_gbs_per_gpu"] / ceil
        # Which host-buffer path is the headline `e2e`: the faster of the two, BOTH measured in this run through the same
        # public call with the copies inside the timed region (`e2e.ascii`, `e2e.host_packed`; one hast_set_option apart).
        # Packing on the host wins when a GPU has ~14 or more host cores to itself (16-core box, N=1: 385 vs 252 M pairs/s)
        # and loses when ranks share the cores (24-core box, N=2, 10 threads per rank: 467 vs 504; profiles/r02_y_host_pack.txt).
        e2e["ascii"] = {k: e2e[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step",
                                            "h2d_gbs_per_gpu", "path") if k in e2e}
        e2e["chosen"] = "ascii"
        if "host_packed" in e2e and e2e["host_packed"]["value"] > e2e["ascii"]["value"]:
            hpk = e2e["host_packed"]
            e2e.update({"value": hpk["value"], "h2d_bytes_per_step": hpk["h2d_bytes_per_step"], "ms_per_step": hpk["ms_per_step"],
                        "h2d_gbs_per_gpu": hpk["h2d_bytes_per_step"] / (hpk["ms_per_step"] * 1e-3) / 1e9, "path": hpk["path"],
                        "chosen": "host_packed"})
        e2e["rule"] = "the faster of e2e.ascii and e2e.host_packed, both measured in this run"
        if ceil:                                   # the ceiling is about bytes over PCIe: restate it for the path chosen
            e2e["ascii"]["frac_of_h2d_ceiling"] = e2e["ascii"]["h2d_gbs_per_gpu"] / ceil
            e2e["frac_of_h2d_ceiling"] = e2e["h2d_gbs_per_gpu"] / ceil

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": WORKLOAD_TEXT[args.workload], "k": spec.k, "read_len": L,
                           "pairs_per_gpu": P, "barcodes": nb, "table_keys": int(info.n_entries),
                           "table_bytes": int(info.bytes), "filter_bytes": int(info.filter_bytes), "sub_batches_per_step": len(batches),
                           "l2": "inputs (%.1f GB of reads per step) are larger than L2; the k-mer table is the "
                                 "workload's own hot state" % (n_reads * (L + 8) / 1e9),
                           "parallelism": f"dp{world}: reads sharded, table replicated, one ncclReduce of counts"},
                "lookups_per_step": int(lookups_step), "lookups_per_s": world * lookups_step / (ms_step * 1e-3),
                "table_build_s": t_table,
                "clocks": clk.summary(), "gpu_launches": int(gpu_launches),
                "parity_checked": bool(parity.get("ok")), "parity": parity,
                "e2e": e2e, "roofline": roofline, "cpu_baseline": None,
                "stats": {"reads_with_n": st["reads_with_n"] // args.steps, "extra_probes": st["extra_probes"] // args.steps,
                          "filter_pass_per_step": st["filter_pass"] // args.steps,
                          "filter_pass_frac": st["filter_pass"] / max(1, st["lookups"]),
                          "filter_loads_per_lookup": st["filter_loads"] / max(1, st["lookups"]),
                          "filter_bytes": int(info.filter_bytes), "kernel": args.kernel}}
    eng.close()
    return line, trio


def h2d_ceiling(world):
    """per-GPU host->device rate with `world` GPUs copying at once, from the committed probe (profiles/h2d_ceiling_r02.json)"""
    try:
        d = json.loads((ROOT / "profiles" / "h2d_ceiling_r02.json").read_text())
        return float(d["per_gpu_gbs"][str(world)])
    except Exception:
        return None


# ----------------------------------------------------------------------------
# configs[2] at its stated shape, strong scaling
# ----------------------------------------------------------------------------
def leg_cfg3(args, R):
    torch = R.torch
    from hast_b200 import synth_stream as ss
    from hast_b200.capi import Engine
    rank, world, dev = R.rank, R.world, R.dev
    spec = ss.stream_config(args.cfg3_workload)
    if args.cfg3_pairs:
        spec.n_pairs = args.cfg3_pairs
    if args.cfg3_barcodes:
        spec.n_barcodes = args.cfg3_barcodes
    L = spec.read_len
    t0 = time.perf_counter()
    trio = ss.StreamTrio(spec, dev)
    torch.cuda.synchronize()
    t_trio = time.perf_counter() - t0
    nb = trio.n_barcodes
    log(f"rank {rank}: cfg3 trio ({spec.genome_len / 1e9:.2f} Gbp) in {t_trio:.1f}s: pat {trio.pat.size} mat {trio.mat.size} k-mers")

    eng = Engine(R.local)
    eng.set_option("kernel", args.kernel)
    eng.set_option("filter_bits_per_key", args.filter_bits)
    eng.set_option("filter_max_bytes", args.filter_max_mib << 20)
    if args.reads_per_tile:
        eng.set_option("reads_per_tile", args.reads_per_tile)
    if args.l2_persist_mib >= 0:
        eng.set_option("l2_persist_bytes", args.l2_persist_mib << 20)
    info, t_table = build_table(eng, spec.k, trio.pat, trio.mat)
    eng.reserve_barcodes(nb)
    R.comm_for(eng)

    # this rank's slice of the pair index space, resident in HBM before the timed region starts
    P_total = spec.n_pairs
    torch.cuda.empty_cache()
    free_b, _ = torch.cuda.mem_get_info()
    per_pair = 2 * L + 8
    fit = int((free_b - (6 << 30)) // per_pair)
    from hast_b200 import dist as hd
    want, _ = hd.strong_slices(P_total, world)
    cap = int(-R.max_over_ranks(-float(max(1, min(want, fit)))))        # the same slicing on every rank
    P_rank, P_total_used = hd.strong_slices(P_total, world, cap)
    reduced_to_fit = P_rank < want
    lo_pair, n_pair = hd.slice_of(rank, P_rank, P_total_used)
    C_pairs = SUB_BATCH_READS // 2
    bases = torch.empty((2 * n_pair, L), dtype=torch.uint8, device=dev)
    bc = torch.empty(2 * n_pair, dtype=torch.int32, device=dev)
    d_off = (torch.arange(2 * C_pairs + 1, dtype=torch.int64, device=dev) * L).to(torch.int32)
    t0 = time.perf_counter()
    batches = []
    for a in range(0, n_pair, C_pairs):
        n = min(C_pairs, n_pair - a)
        bp, cp = bases.data_ptr() + 2 * a * L, bc.data_ptr() + 8 * a
        trio.gen_pairs_into(lo_pair + a, n, bp, cp)
        batches.append((bp, 2 * n * L, d_off.data_ptr(), cp, 2 * n))
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    log(f"rank {rank}: cfg3 {n_pair} pairs ({2 * n_pair * L / 1e9:.1f} GB of bases) generated on the device in {t_gen:.1f}s, "
        f"{len(batches)} launches per step, table {info.bytes / 2**20:.0f} MiB, {nb} barcodes")

    def step(collect=True):
        for b in batches:
            eng.submit_batch_device(*b)
        if collect:
            return eng.finish(nb, want_counts=(rank == 0), pinned=True)

    steps = args.cfg3_steps
    for _ in range(2):
        eng.reset_counts()
        step()
    eng.reset_counts()
    launches0 = eng.stats()["kernel_launches"]
    R.barrier()
    with ClockSampler(R.local) as clk:
        eng.timer_start()
        for _ in range(steps):
            counts = step()
        ms_total = eng.timer_stop()
        R.barrier()
    st = eng.stats()
    ms_step = R.max_over_ranks(ms_total / steps)
    reduce_ms = R.max_over_ranks(st["finish_reduce_us"] / 1e3 / steps)
    d2h_ms = st["finish_d2h_us"] / 1e3 / steps
    lookups_rank = st["lookups"] // steps
    lookups_total = R.sum_over_ranks([lookups_rank])[0]
    gpu_launches = st["kernel_launches"] - launches0

    # ---- parity, in this run -------------------------------------------------------------------
    loc = local_count_sums(R, eng, nb)
    tot = R.sum_over_ranks(loc)
    parity = None
    if rank == 0:
        red = counts.astype(np.int64).sum(0)
        sums_ok = [int(red[0]), int(red[1])] == tot
        t0 = time.perf_counter()
        rng = np.random.default_rng(11)
        ids = np.sort(rng.choice(spec.n_barcodes, size=min(64, spec.n_barcodes), replace=False))
        idx = trio.pairs_of_barcodes(ids, 0, P_total_used)           # over ALL ranks' slices
        sb, s_bc = trio.gen_pairs_idx(idx)
        o = oracle_for(trio.pat, trio.mat, spec.k)
        s_off = np.arange(sb.shape[0] + 1, dtype=np.uint64) * np.uint64(L)
        want, _ = o.classify_batch(sb.reshape(-1), s_off, s_bc, nb, nthreads=os.cpu_count() or 4)
        same = bool((counts[ids].astype(np.int64) == steps * want[ids].astype(np.int64)).all())
        sizes_ok = (int(info.size[0]), int(info.size[1])) == (o.set_size(0), o.set_size(1))
        by_rank = np.bincount(np.minimum(idx // max(P_rank, 1), world - 1), minlength=world)
        o.close()
        parity = {"ok": bool(sums_ok and same and sizes_ok and want[ids].sum() > 0),
                  "sum_of_partials_equals_reduced": sums_ok, "reduced_sums": [int(red[0]), int(red[1])],
                  "oracle": "oracle/hast_oracle.c ho_classify_batch, full k-mer lists (62 M keys at cfg3)",
                  "oracle_barcodes": int(ids.size), "oracle_pairs": int(idx.size),
                  "oracle_pairs_by_rank": [int(x) for x in by_rank], "oracle_hits": int(want[ids].sum()),
                  "oracle_counts_identical": same, "oracle_set_sizes_identical": sizes_ok,
                  "seconds": time.perf_counter() - t0}
        if not parity["ok"]:
            raise SystemExit(f"PARITY FAILURE in the cfg3 leg: {parity}")

    eng.reset_counts()
    R.barrier()
    eng.timer_start()
    step(collect=False)
    ms_kernel = eng.timer_stop()
    eng.sync()
    out = None
    if rank == 0:
        peak, peak_source = load_peaks()
        gather_tbl = eng.gather_roofline(1 << 28, info.bytes)
        roof = roofline_of(lookups_rank / max(1, len(batches)), ms_kernel / max(1, len(batches)), peak, peak_source,
                           "fused_kernel_dram_bytes_per_launch_cfg3", None, gather_tbl, info)
        roof["frac_of_random_gather"] = roof["achieved"] / gather_tbl if gather_tbl else None
        out = {"workload": "configs[2]: synthetic %.1f Gbp diploid trio (het %.5f per haplotype), k=%d, %d parent-unique k-mers "
                           "(exact set difference), %d read pairs over %d barcodes, strong scaling over %d GPU(s)"
                           % (spec.genome_len / 1e9, spec.het, spec.k, int(info.n_entries), P_total_used, spec.n_barcodes, world),
               "value": P_total_used / (ms_step * 1e-3), "unit": UNIT, "scaling": "strong", "n_gpus": world,
               "steps": steps, "warmup": 2, "ms_per_step": ms_step,
               "kernel_ms": ms_kernel, "reduce_ms": reduce_ms, "d2h_ms": d2h_ms,
               "reduce_share_of_step": reduce_ms / ms_step, "d2h_share_of_step": d2h_ms / ms_step,
               "reduce_bytes": int(nb * 8) if world > 1 else 0, "d2h_bytes": int(nb * 8),
               "pairs_total": int(P_total_used), "pairs_per_gpu": int(P_rank), "pairs_stated": int(P_total),
               "reduced_to_fit_hbm": reduced_to_fit, "barcodes": int(nb),
               "table_keys": int(info.n_entries), "table_bytes": int(info.bytes), "filter_bytes": int(info.filter_bytes),
               "set_sizes": [int(info.size[0]), int(info.size[1])],
               "launches_per_step": len(batches), "gpu_launches": int(gpu_launches),
               "lookups_per_s": lookups_total / (ms_step * 1e-3),
               "hbm_bytes_resident_per_gpu": int(2 * n_pair * (L + 4) + info.bytes + info.filter_bytes + nb * 8),
               "generate_s": {"trio": t_trio, "table": t_table, "reads": t_gen},
               "clocks": clk.summary(), "parity": bool(parity["ok"]), "parity_detail": parity, "roofline": roof}
    eng.close()
    del bases, bc, trio
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------
# the drop-in process against the reference process
# ----------------------------------------------------------------------------
def leg_cli(args, trio, runner, workdir: Path):
    """bin/classify vs oracle/_ref/classify_O2 on the same plain and gzip FASTQ files (like for like: both sides open
    the files, decompress, parse, classify and print the table)."""
    exe = ROOT / "bin" / "classify"
    cores = os.cpu_count() or 8
    threads = max(4, cores - 2)
    pairs = runner.sample
    out = {"pairs": int(pairs), "host_cores": cores, "parser_threads": threads, "gpus": 1,
           "fastq_bytes": os.path.getsize(runner.r1) + os.path.getsize(runner.r2),
           "gz_bytes": sum(os.path.getsize(p) for p in runner.gz), "gz": "one gzip member per file, level 6",
           "legs": {}}

    def ours(name, reads):
        stats = workdir / f"{name}.json"
        cmd = [str(exe), "--hap0", runner.pat, "--hap1", runner.mat, "--weight0", "1.04", "--thread", str(threads),
               "--gpus", "1", "--stats-json", str(stats)]
        for r in reads:
            cmd += ["--read", r]
        best = None
        for _ in range(3):
            t = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True)
            dt = time.perf_counter() - t
            if r.returncode != 0:
                raise SystemExit(f"bin/classify failed: {r.stderr[-800:]}")
            s = json.loads(stats.read_text())
            if best is None or dt < best[0]:
                best = (dt, s, r.stdout)
        dt, s, table = best
        out["legs"][name] = {"wall_s": dt, "pairs_per_s": pairs / dt, "stream_s": s["t_reads_s"],
                             "pairs_per_s_stream": s["pairs_per_s_stream"], "t_table_s": s["t_table_s"],
                             "t_print_s": s["t_print_s"], "text_GBps_stream": s["fastq_text_bytes"] / s["t_reads_s"] / 1e9}
        return table

    t_plain = ours("plain", [runner.r1, runner.r2])
    t_gz = ours("gz", list(runner.gz))
    ref_plain = runner.table
    dt_gz, ref_gz = runner.timed(runner.cmd(runner.threads, list(runner.gz)))
    out["legs"]["reference_plain"] = {"wall_s": runner.last_wall, "pairs_per_s": pairs / runner.last_wall, "threads": runner.threads}
    out["legs"]["reference_gz"] = {"wall_s": dt_gz, "pairs_per_s": pairs / dt_gz, "threads": runner.threads,
                                   "binary": "oracle/_ref/classify_O2 (untouched reference sources, -O2)"}
    out["tables_identical"] = bool(t_plain == t_gz == ref_plain == ref_gz)
    out["table_lines"] = t_plain.count(b"\n")
    out["speedup_wall"] = {"plain": runner.last_wall / out["legs"]["plain"]["wall_s"], "gz": dt_gz / out["legs"]["gz"]["wall_s"]}
    net = lambda w: max(w - runner.t_load, 1e-6)
    out["speedup_streaming"] = {"plain": out["legs"]["plain"]["pairs_per_s_stream"] / (pairs / net(runner.last_wall)),
                                "gz": out["legs"]["gz"]["pairs_per_s_stream"] / (pairs / net(dt_gz))}
    if not out["tables_identical"]:
        raise SystemExit("CLI leg: bin/classify and the reference binary printed different tables")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hast_b200", choices=["hast_b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOAD_TEXT))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the configs[2] strong-scaling leg")
    ap.add_argument("--no-cli", action="store_true", help="skip the bin/classify vs reference process leg")
    ap.add_argument("--cfg3-workload", default="cfg3", help="synth_stream config of the cfg3 leg (cfg3 | cfg3_100m | stream_small)")
    ap.add_argument("--cfg3-pairs", type=int, default=0, help="override the leg's total pair count")
    ap.add_argument("--cfg3-barcodes", type=int, default=0)
    ap.add_argument("--cfg3-steps", type=int, default=3)
    ap.add_argument("--table-scale", type=float, default=1.0, help="expected_keys multiplier (sparser table)")
    ap.add_argument("--kernel", type=int, default=3, choices=[0, 1, 2, 3, 4],
                    help="3 = pre-filtered fused kernel, filter word chosen by the k-mer's minimizer (default), "
                         "1 = filter word chosen by a hash of the k-mer, 2 = as 1 with TMA-staged reads, "
                         "0 = direct table probe per position")
    ap.add_argument("--filter-bits", type=int, default=16, help="pre-filter bits per key")
    ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (32/64/128), 0 = leave")
    ap.add_argument("--filter-max-mib", type=int, default=64, help="pre-filter size cap (MiB)")
    ap.add_argument("--reads-per-tile", type=int, default=0, help="fused kernel: reads per tile (0 = what fills one pass)")
    ap.add_argument("--l2-persist-mib", type=int, default=-1, help="tuning: L2 set-aside for persisting accesses (MiB); -1 = leave")
    ap.add_argument("--host-pack-threads", type=int, default=-1, help="e2e.host_packed leg: host threads that pack (-1 = cores/ranks - 2, 0 = skip the leg)")
    ap.add_argument("--only-cfg3", action="store_true", help="tuning: run the cfg3 leg alone and print its object")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "hast_b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    R = Runtime()
    if args.only_cfg3:
        c3 = leg_cfg3(args, R)
        if R.rank == 0:
            emit(c3)
        R.close()
        return 0
    line, trio = leg_cfg2(args, R)

    # ---- CPU baseline and the process-level comparison beside it (rank 0, N=1 only) -----------
    if R.rank == 0 and R.world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        want_cli = not args.no_cli and (ROOT / "oracle" / "_ref" / "classify_O2").exists() and (ROOT / "bin" / "classify").exists()
        sample = min(trio.spec.n_pairs, max(100_000, min(4_000_000, 250_000 * cores)))
        with tempfile.TemporaryDirectory(prefix="hast_cpu_", dir=os.environ.get("TMPDIR", "/tmp")) as d:
            runner = RefRunner(trio, sample, Path(d), gz_too=want_cli)
            line["cpu_baseline"] = runner.run()
            if want_cli:
                runner.last_wall = line["cpu_baseline"]["wall_s"]
                line["cli"] = leg_cli(args, trio, runner, Path(d))
    del trio
    R.torch.cuda.empty_cache()

    if not args.no_cfg3:
        c3 = leg_cfg3(args, R)
        if R.rank == 0:
            line["cfg3"] = c3
            line["roofline"]["hbm_resident"] = c3["roofline"]        # the 1 GiB-table launch: the HBM-resident data point
    if R.rank == 0:
        emit(line)
    R.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
