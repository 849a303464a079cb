#!/usr/bin/env python
"""bench.py -- stLFR read pairs classified per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

A *step* is one pass of the hot path (pack -> canonical k-mers -> table probe ->
per-barcode reduce, then the count collection) over the whole synthetic workload
of this rank.  At N=1 the workload is BASELINE.json configs[1]: a synthetic
100 Mbp diploid trio, k=21 parent-unique k-mers, 20 M stLFR 100 bp read pairs
over 500 k barcodes.  At N>1 every rank classifies its own 20 M pairs of the
same trio (weak scaling), the k-mer table is replicated per GPU and the
per-barcode partial counts are summed with ONE ncclReduce inside hast_finish.

value        device-resident inputs, CUDA-event time on the launching stream, max over ranks
e2e          the same step through hast_submit_batch with HOST (pinned) buffers, H2D copies and
             the D2H read of the counts inside the timed region
roofline     fused kernel: 32 B (one table sector) per k-mer lookup / kernel time  vs measured HBM peak
cpu_baseline the UNTOUCHED reference binary (oracle/_ref/classify_O2, all host threads) on a bounded
             sample of the same workload, k-mer load time subtracted
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
class RefRunner:
    """The UNTOUCHED reference binary (oracle/_ref/classify_O2, built from /root/reference sources by
    oracle/Makefile) on the first sample_pairs pairs of the workload, written as FASTQ.  pairs/s is net of
    the k-mer load time, measured once with a one-read FASTQ.  Falls back to the
    plain-C oracle port when the reference binary is absent."""

    def __init__(self, trio, sample_pairs: int, workdir: Path):
        self.trio, self.sample = trio, min(sample_pairs, trio.spec.n_pairs)
        self.ref = ROOT / "oracle" / "_ref" / "classify_O2"
        self.cores = os.cpu_count() or 1
        self.t_load = None
        if self.ref.exists():
            self.pat, self.mat = trio.write_kmer_lists(workdir)
            self.r1, self.r2 = trio.write_fastq(workdir, gz=False, lo=0, hi=self.sample, stem="sample")
            self.one, _ = trio.write_fastq(workdir, gz=False, lo=0, hi=1, stem="one")
        else:
            sys.path.insert(0, str(ROOT / "tests"))
            import oracle as orc
            self.o = orc.Oracle()
            self.o.load_kmers(trio.kmer_text(0), 0)
            self.o.load_kmers(trio.kmer_text(1), 1)
            self.o.init_adaptor()
            self.batch = trio.batch(0, self.sample)
        self.threads = None

    def _cmd(self, threads, reads):
        c = [str(self.ref), "--hap0", self.pat, "--hap1", self.mat, "--weight0", "1.04", "--thread", str(threads)]
        for r in reads:
            c += ["--read", r]
        return c

    def _timed(self, cmd):
        t = time.perf_counter()
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return time.perf_counter() - t

    def run(self):
        if not self.ref.exists():
            bases, off, bc = self.batch
            t = time.perf_counter()
            self.o.classify_batch(bases, off.astype(np.uint64), bc, self.trio.n_barcodes, nthreads=self.cores)
            net = time.perf_counter() - t
            return {"value": self.sample / net, "unit": UNIT, "cores": self.cores, "kind": "port",
                    "sample": f"first {self.sample} pairs, oracle/hast_oracle.c ho_classify_batch, "
                              f"{self.cores} threads, in memory"}
        if self.t_load is None:
            self.t_load = self._timed(self._cmd(8, [self.one]))
        if self.threads is None:             # the reference stops scaling early (one reader thread,
            best = None                      # classify.cpp:257-269): sweep and keep the best, BASELINE.md 3.2
            for t in sorted({min(8, self.cores), min(16, self.cores), min(32, self.cores), self.cores}):
                dt = self._timed(self._cmd(t, [self.r1, self.r2]))
                log(f"reference --thread {t}: {dt:.2f}s")
                if best is None or dt < best[1]:
                    best = (t, dt)
            self.threads, t_all = best
        else:
            t_all = self._timed(self._cmd(self.threads, [self.r1, self.r2]))
        net = max(t_all - self.t_load, 1e-6)
        return {"value": self.sample / net, "unit": UNIT, "cores": self.threads, "host_cores": self.cores,
                "kind": "reference",
                "sample": f"first {self.sample} pairs of the workload as plain FASTQ; oracle/_ref/classify_O2 "
                          f"(untouched reference sources, -O2) --thread {self.threads} (best of sweep); "
                          f"wall {t_all:.2f}s minus k-mer load {self.t_load:.2f}s",
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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hast_b200", choices=["hast_b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOAD_TEXT))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--table-scale", type=float, default=1.0, help="expected_keys multiplier (sparser table)")
    ap.add_argument("--kernel", type=int, default=3, choices=[0, 1, 2, 3, 4],
                    help="3 = pre-filtered fused kernel, filter word chosen by the k-mer's minimizer (default), "
                         "1 = filter word chosen by a hash of the k-mer, 2 = as 1 with TMA-staged reads, "
                         "0 = direct table probe per position")
    ap.add_argument("--filter-bits", type=int, default=16, help="pre-filter bits per key")
    ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (32/64/128), 0 = leave")
    ap.add_argument("--filter-max-mib", type=int, default=64, help="pre-filter size cap (MiB)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "hast_b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from hast_b200 import synth
    from hast_b200.capi import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the classification path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload ------------------------------------------------------------
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

    eng = Engine(local)
    eng.set_option("kernel", args.kernel)
    eng.set_option("filter_bits_per_key", args.filter_bits)
    eng.set_option("filter_max_bytes", args.filter_max_mib << 20)
    if args.l2_fetch:
        eng.set_option("l2_fetch_granularity", args.l2_fetch)
    n_keys = trio.pat.size + trio.mat.size
    eng.table_begin(spec.k, int(n_keys * args.table_scale))
    t0 = time.perf_counter()
    eng.table_add_packed(trio.pat, 0)
    eng.table_add_packed(trio.mat, 1)
    eng.table_erase_seq(b"CTGTCTCTTATACACATCTTAGGAAGACAAGCACTGACGACATGA")
    eng.table_erase_seq(b"TCTGCTGAGTCGAGAACGTCTCTGTGAGCCAAGGAGTTGCTCTGG")
    info = eng.table_info()
    t_table = time.perf_counter() - t0
    log(f"rank {rank}: table {info.bytes / 2**20:.0f} MiB + pre-filter {info.filter_bytes / 2**20:.1f} MiB, {info.n_entries} entries, "
        f"{info.n_overflow_buckets} overflow buckets, {info.n_displaced} displaced, built in {t_table:.2f}s")
    eng.reserve_barcodes(nb)
    if world > 1:
        uid = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init_rank(world, rank, uid[0])

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
    barrier()
    with ClockSampler(local) as clk:
        eng.timer_start()
        for _ in range(args.steps):
            counts = step_device()
        ms_total = eng.timer_stop()
        barrier()
    st = eng.stats()
    gpu_launches = st["kernel_launches"] - launches0
    ms_step = max_over_ranks(ms_total / args.steps)
    lookups_step = st["lookups"] // args.steps
    value = world * P / (ms_step * 1e-3)

    # kernel-only duration for the roofline (no finish / reduce / D2H in the region)
    eng.reset_counts()
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        step_device(collect=False)
    ms_kernel = eng.timer_stop() / args.steps
    eng.sync()
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = lookups_step * ALG_BYTES_PER_LOOKUP / (ms_kernel * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("fused_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    gather = eng.gather_roofline(1 << 28, 4 << 30) if rank == 0 else None
    gather_tbl = eng.gather_roofline(1 << 28, info.bytes) if rank == 0 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "classify_kernel" if args.kernel >= 1 else "tile_kernel<MODE_CLASSIFY>", "launches_per_step": len(batches),
                "ms_per_launch": ms_kernel / len(batches),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
                "lookups_per_s": lookups_step / (ms_kernel * 1e-3),
                "random_gather_gbs_4GiB": gather, "random_gather_gbs_table_span": gather_tbl,
                "frac_of_random_gather_4GiB": (achieved / gather) if gather else None,
                "note": "table of %d MiB vs 126 MB L2: probes are partly L2 hits, so achieved may exceed the HBM "
                        "random-gather figure" % (info.bytes >> 20)}

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
        barrier()
        t = time.perf_counter()
        for _ in range(args.steps):
            c_h = step_host()
        eng.sync()
        dt = max_over_ranks((time.perf_counter() - t) / args.steps)
        barrier()
        if rank == 0:                      # both legs accumulated `steps` identical passes
            assert (c_h == counts).all(), "host-buffer path and device-resident path disagree"
        h2d = n_reads * L + sum(b[4] + 1 for b in hb) * 4 + n_reads * 4
        e2e = {"value": world * P / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(nb * 8), "ms_per_step": dt * 1e3,
               "h2d_gbs_per_gpu": h2d / dt / 1e9,        # against ~55 GB/s of a PCIe Gen5 x16 link: this leg is PCIe-bound
               "path": "hast_submit_batch (pinned host buffers, double-buffered cudaMemcpyAsync) + hast_finish"}
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
            barrier()
            t = time.perf_counter()
            for _ in range(args.steps):
                c_p = step_packed()
            eng.sync()
            dtp = max_over_ranks((time.perf_counter() - t) / args.steps)
            barrier()
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
    # ---- CPU baseline beside it (rank 0, N=1 only) ------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = min(P, max(100_000, min(2_000_000, 62_500 * cores)))
        with tempfile.TemporaryDirectory(prefix="hast_cpu_") as d:
            cpu = RefRunner(trio, sample, Path(d)).run()
        # parity spot check of the timed path on the same sample happens in tests/; here only timing

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
                "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
                "stats": {"reads_with_n": st["reads_with_n"] // args.steps, "extra_probes": st["extra_probes"] // args.steps,
                          "filter_pass_per_step": st["filter_pass"] // args.steps,
                          "filter_pass_frac": st["filter_pass"] / max(1, st["lookups"]),
                          "filter_loads_per_lookup": st["filter_loads"] / max(1, st["lookups"]),
                          "filter_bytes": int(info.filter_bytes), "kernel": args.kernel}}
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
