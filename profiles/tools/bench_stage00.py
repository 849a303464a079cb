#!/usr/bin/env python
"""Throughput of the stage-00 count table (hast_kc_*, csrc/kcount.cuh) on one B200.

    python profiles/tools/bench_stage00.py [--genome 5000000] [--coverage 30] [--steps 3] > out.json

Workload: whole-genome shotgun reads of both parents of the synthetic trio (hast_b200/synth.py
parent_reads: 100 bp, 0.3 % errors, 1 % reads with an N), k = 21.
  value      k-mer windows counted per second, reads resident in HBM, CUDA-event time of kc_count_kernel
  e2e        bin/build_unshared_kmers --auto_bounds on the same reads as plain FASTQ files: wall clock of the
             whole stage (read, parse, H2D, count, histogram, bounds, select, sort, write the lists)
  roofline   one random 16-byte slot read-modify-write per window; compared with the measured random
             32-byte-sector gather rate over the table's span (hast_gather_roofline)
"""
import argparse
import json
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--l2-fetch", type=int, default=0)
    args = ap.parse_args()
    import torch
    from hast_b200 import synth
    from hast_b200.capi import Engine

    k, L = 21, 100
    spec = synth.TrioSpec(genome_len=args.genome, het=0.001, k=k)
    t0 = time.perf_counter()
    reads = synth.parent_reads(spec, args.coverage, read_len=L)
    print(f"[stage00] reads generated in {time.perf_counter() - t0:.1f}s", file=sys.stderr)
    dev = "cuda:0"
    d = {n: torch.from_numpy(reads[n]).to(dev) for n in ("paternal", "maternal")}
    n_reads = {n: reads[n].shape[0] for n in d}
    SUB = 2_000_000
    d_off = (torch.arange(SUB + 1, dtype=torch.int64, device=dev) * L).to(torch.int32)
    eng = Engine(0)
    if args.l2_fetch:
        eng.set_option("l2_fetch_granularity", args.l2_fetch)
    expected = int(2.2 * args.genome * (1 + args.coverage * 0.003 * k))       # genomic + error k-mers, both parents
    out = {}

    def one_pass():
        for p, n in ((0, "paternal"), (1, "maternal")):
            for lo in range(0, n_reads[n], SUB):
                m = min(SUB, n_reads[n] - lo)
                eng.kc_add_device(d[n].data_ptr() + lo * L, m * L, d_off.data_ptr(), m, p)

    ms = []
    for step in range(args.steps + 1):
        eng.kc_begin(k, expected)
        eng.sync()
        eng.timer_start()
        one_pass()
        t = eng.timer_stop()
        if step:
            ms.append(t)
    ki = eng.kc_info()
    assert ki.table_full == 0
    windows = ki.windows
    t_histo = time.perf_counter()
    h = eng.kc_histo(0, 10000)
    t_histo = time.perf_counter() - t_histo
    t_sel = time.perf_counter()
    sel = eng.kc_select(0, 9, 90)
    t_sel = time.perf_counter() - t_sel
    gather = eng.gather_roofline(1 << 27, ki.bytes)
    best = min(ms)
    out = {"metric": "stage-00 k-mer windows counted/s", "value": windows / (best * 1e-3), "unit": "windows/s",
           "ms_per_pass": best, "ms_all": ms, "windows": int(windows), "reads": int(sum(n_reads.values())),
           "config": {"workload": f"synthetic parents of a {args.genome} bp trio, {args.coverage}x each, 100 bp reads, k=21",
                      "table_bytes": int(ki.bytes), "slots": int(ki.n_slots), "occupied": int(ki.occupied),
                      "load": ki.occupied / ki.n_slots},
           "distinct": [int(ki.distinct[0]), int(ki.distinct[1])], "both": int(ki.both),
           "histo_s": t_histo, "select_sort_s": t_sel, "selected": int(sel.size),
           "roofline": {"bound": "hbm", "unit": "G slot updates/s", "achieved": windows / (best * 1e-3) / 1e9,
                        "random_gather_sectors_per_s_over_table_span": gather / 32.0,
                        "frac_of_random_gather": windows / (best * 1e-3) / 1e9 / (gather / 32.0),
                        "note": "one random 16-byte slot RMW (key compare + atomicAdd) per window"},
           "gpu_launches": eng.stats()["kernel_launches"]}
    eng.kc_end()
    eng.close()
    if not args.no_e2e:
        with tempfile.TemporaryDirectory(prefix="hast_s00_") as td:
            td = Path(td)
            pat = synth.write_reads_fastq(td / "pat.fq", reads["paternal"])
            mat = synth.write_reads_fastq(td / "mat.fq", reads["maternal"])
            run = td / "run"
            run.mkdir()
            t = time.perf_counter()
            r = subprocess.run([str(ROOT / "bin" / "build_unshared_kmers"), "--paternal", pat, "--maternal", mat, "--mer", "21",
                                "--auto_bounds", "--thread", "8", "--gpus", "1", "--stats-json", str(run / "stats.json")],
                               cwd=run, capture_output=True, text=True)
            dt = time.perf_counter() - t
            assert r.returncode == 0, r.stdout[-800:]
            st = json.loads((run / "stats.json").read_text())
            out["e2e"] = {"wall_s": dt, "value": windows / dt, "unit": "windows/s", "stats": st,
                          "text_bytes": st["text_bytes"], "path": "bin/build_unshared_kmers --auto_bounds, plain FASTQ, 1 GPU"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
