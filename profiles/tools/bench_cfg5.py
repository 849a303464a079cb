#!/usr/bin/env python
"""BASELINE.json configs[4] at its stated shape, through the drop-in process.

    python profiles/tools/bench_cfg5.py [--pairs 200000000] [--barcodes 50000000] [--gpus-list 1,8] > out.json

Workload (hast_b200/synth_stream.py "cfg5"): the human-scale trio of configs[2] (3.1 Gbp, ~62 M parent-unique
21-mers), 50 M barcode NAMES with Zipf(1.2) reads per barcode, >= 200 M read pairs, written as two gzip FASTQ files
(one member each).  Legs: bin/classify --gpus 1 and --gpus N on the full files (whole-process wall, streaming rate,
k-mer list load, table print; the two tables must be byte-identical), then the reference binary on a
BARCODE-COMPLETE subsample (every pair of every barcode whose id is 7 mod 32): its table
must be, line for line, the rows of those barcodes in bin/classify's table.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def log(*a):
    print("[cfg5]", *a, file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=200_000_000)
    ap.add_argument("--barcodes", type=int, default=50_000_000)
    ap.add_argument("--config", default="cfg5")
    ap.add_argument("--gpus-list", default="1,8")
    ap.add_argument("--gz-level", type=int, default=6)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--workdir", default=os.environ.get("TMPDIR", "/tmp"))
    args = ap.parse_args()
    import torch
    from hast_b200 import synth_stream as ss
    cores = os.cpu_count() or 8
    threads = args.threads or max(4, cores - 2)
    spec = ss.stream_config(args.config)
    spec.n_pairs, spec.n_barcodes = args.pairs, args.barcodes
    t0 = time.perf_counter()
    trio = ss.StreamTrio(spec, "cuda:0")
    torch.cuda.synchronize()
    out = {"workload": f"configs[4]: {spec.genome_len / 1e9:.1f} Gbp trio, k={spec.k}, {trio.pat.size + trio.mat.size} parent-unique "
                       f"k-mers, {spec.n_pairs} read pairs, {spec.n_barcodes} barcode names, Zipf({spec.zipf_alpha}) reads per barcode, "
                       f"gzip level {args.gz_level} (one member per file)",
           "host_cores": cores, "parser_threads": threads, "generate_s": {"trio": time.perf_counter() - t0}, "legs": {}}
    log(f"trio in {out['generate_s']['trio']:.1f}s")
    with tempfile.TemporaryDirectory(prefix="hast_cfg5_", dir=args.workdir) as d:
        d = Path(d)
        t0 = time.perf_counter()
        pat, mat = trio.write_kmer_lists(d)
        names = trio.barcode_name_blob()
        out["generate_s"]["lists_and_names"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        gz = trio.write_fastq(d, gz=args.gz_level, names=names)
        out["generate_s"]["fastq_gz"] = time.perf_counter() - t0
        out["gz_bytes"] = sum(os.path.getsize(p) for p in gz)
        log(f"lists+names {out['generate_s']['lists_and_names']:.1f}s, fastq.gz {out['generate_s']['fastq_gz']:.1f}s, {out['gz_bytes'] / 1e9:.2f} GB")
        # the GPUs are needed by the processes under test from here on
        exe = str(ROOT / "bin" / "classify")
        tables = {}
        for g in [int(x) for x in args.gpus_list.split(",") if x]:
            if g > torch.cuda.device_count():
                continue
            stats = d / f"stats_{g}.json"
            table = d / f"table_{g}.txt"
            cmd = [exe, "--hap0", pat, "--hap1", mat, "--weight0", "1.04", "--thread", str(threads), "--gpus", str(g),
                   "--stats-json", str(stats), "--read", gz[0], "--read", gz[1]]
            t = time.perf_counter()
            with open(table, "wb") as f:
                r = subprocess.run(cmd, stdout=f, stderr=subprocess.PIPE)
            dt = time.perf_counter() - t
            assert r.returncode == 0, r.stderr[-1500:]
            s = json.loads(stats.read_text())
            h = hashlib.md5()
            with open(table, "rb") as f:
                for blk in iter(lambda: f.read(1 << 24), b""):
                    h.update(blk)
            tables[g] = h.hexdigest()
            out["legs"][f"gpus_{g}"] = {"wall_s": dt, "pairs_per_s": spec.n_pairs / dt, "stream_s": s["t_reads_s"],
                                        "pairs_per_s_stream": s["pairs_per_s_stream"], "t_table_s": s["t_table_s"],
                                        "t_finish_s": s["t_finish_s"], "t_print_s": s["t_print_s"], "barcodes_seen": s["barcodes"],
                                        "text_GBps_stream": s["fastq_text_bytes"] / s["t_reads_s"] / 1e9, "table_md5": tables[g],
                                        "table_bytes": os.path.getsize(table)}
            log(f"--gpus {g}: wall {dt:.1f}s, stream {s['t_reads_s']:.1f}s = {s['pairs_per_s_stream'] / 1e6:.2f} M pairs/s, "
                f"table load {s['t_table_s']:.1f}s, print {s['t_print_s']:.1f}s, {s['barcodes']} barcodes seen")
        out["tables_identical_across_gpu_counts"] = len(set(tables.values())) == 1
        ref = ROOT / "oracle" / "_ref" / "classify_O2"
        if ref.exists() and not args.no_reference and tables:
            t0 = time.perf_counter()
            ids = np.arange(7, spec.n_barcodes, 32, dtype=np.int64)
            idx = trio.pairs_of_barcodes(ids)
            sub = trio.write_fastq(d / "sub", pair_idx=np.sort(idx), gz=False, stem="sub", names=names)
            out["subsample"] = {"rule": "every pair of every barcode with id = 7 (mod 32)", "barcodes": int(ids.size),
                                "pairs": int(idx.size), "fraction_of_pairs": idx.size / spec.n_pairs,
                                "write_s": time.perf_counter() - t0}
            rt = min(16, cores)
            t = time.perf_counter()
            r = subprocess.run([str(ref), "--hap0", pat, "--hap1", mat, "--weight0", "1.04", "--thread", str(rt),
                                "--read", sub[0], "--read", sub[1]], capture_output=True)
            dt = time.perf_counter() - t
            assert r.returncode == 0, r.stderr[-800:]
            ref_lines = r.stdout.splitlines()
            # the reference's rows must be a subsequence of ours (same order: both sorted bytewise) and every row of
            # ours that names a subsample barcode must be among them
            g0 = sorted(tables)[0]
            j = 0
            with open(d / f"table_{g0}.txt", "rb") as f:
                for line in f:
                    if j < len(ref_lines) and line.rstrip(b"\n") == ref_lines[j]:
                        j += 1
            out["subsample"].update({"reference_wall_s": dt, "reference_threads": rt, "reference_pairs_per_s": idx.size / dt,
                                     "reference_rows": len(ref_lines), "rows_found_identical_in_our_table": j,
                                     "identical": j == len(ref_lines) and j > 0})
            log(f"reference on the subsample: {dt:.1f}s, {len(ref_lines)} rows, {j} found byte-identical in bin/classify's table")
            assert out["subsample"]["identical"], "reference rows differ from bin/classify's table"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
