#!/usr/bin/env python
"""Stand-alone kernels and the k sweep (BASELINE.json configs[3]) on one B200.

    python profiles/tools/bench_kernels.py [--sweep] [--reads 4000000] > out.json

  K2  tile_kernel<MODE_EXTRACT>  (chopRead2Kmer, kmer.h:169-194): L bytes in, 8 bytes out per k-mer position;
      streaming kernel, compared with the measured HBM copy peak (MEASURED_PEAKS.json)
  K3  lookup_kernel (unordered_set::find x2, classify.cpp:195-202): 8 bytes in, one 32-byte table sector, 1 byte out
      per lookup; random-access kernel, compared with the random 32-byte gather rate over the same table span
  --sweep: k = 17 / 21 / 25 / 31 on the 100 Mbp trio: keys, table and filter size, fused-kernel lookups/s and
      algorithmic GB/s (32 B per lookup), K3 GB/s
All times are CUDA events on the launching stream (hast_timer_*), after a warm-up launch.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def measure(k, n_reads, torch, synth, Engine, peaks, decoys=0, fused=True):
    spec = synth.config("cfg2")
    spec.k = k
    spec.n_pairs = n_reads // 2
    spec.decoy_kmers = decoys
    t = synth.make_trio(spec, device="cuda:0", keep_reads_on_device=True)
    L, P = spec.read_len, spec.n_pairs
    n = 2 * P
    d_bases = torch.as_strided(t.r1, (n * L,), (1,))
    d_off = (torch.arange(n + 1, dtype=torch.int64, device="cuda:0") * L).to(torch.int32)
    d_bc = torch.from_numpy(np.concatenate([t.pair_bc, t.pair_bc]).astype(np.int32)).to("cuda:0")
    e = Engine(0)
    e.table_begin(k, t.pat.size + t.mat.size)
    e.table_add_packed(t.pat, 0)
    e.table_add_packed(t.mat, 1)
    info = e.table_info()
    out = {"k": k, "reads": n, "keys": int(info.n_entries), "table_bytes": int(info.bytes), "filter_bytes": int(info.filter_bytes)}
    # K2
    d_km = torch.full((n * L,), -1, dtype=torch.int64, device="cuda:0")
    d_hasn = torch.zeros(n, dtype=torch.uint8, device="cuda:0")
    ms = []
    for i in range(4):
        e.sync()
        e.timer_start()
        e.extract_kmers_device(d_bases.data_ptr(), n * L, d_off.data_ptr(), n, d_km.data_ptr(), d_hasn.data_ptr())
        ms.append(e.timer_stop())
    valid = d_km != -1
    n_pos = int(valid.sum())
    t2 = min(ms[1:])
    k2_bytes = n * (L + 4) + n_pos * 8
    out["k2_extract"] = {"ms": t2, "positions": n_pos, "algorithmic_bytes": k2_bytes, "GBps": k2_bytes / t2 / 1e6,
                         "frac_of_hbm_copy_peak": k2_bytes / t2 / 1e6 / peaks["hbm_gbs"]}
    # K3 on the valid k-mers of non-N reads
    keep = valid & (d_hasn.repeat_interleave(L) == 0)
    km = d_km[keep].contiguous()
    del d_km, valid, keep
    d_tags = torch.zeros(km.numel(), dtype=torch.uint8, device="cuda:0")
    ms = []
    for i in range(4):
        e.sync()
        e.timer_start()
        e.lookup_device(km.data_ptr(), km.numel(), d_tags.data_ptr())
        ms.append(e.timer_stop())
    t3 = min(ms[1:])
    gather = e.gather_roofline(1 << 28, info.bytes)
    out["k3_lookup"] = {"ms": t3, "lookups": int(km.numel()), "lookups_per_s": km.numel() / t3 * 1e3,
                        "table_GBps": km.numel() * 32 / t3 / 1e6, "total_GBps": km.numel() * 41 / t3 / 1e6,
                        "random_gather_GBps_same_span": gather, "frac_of_random_gather": km.numel() * 32 / t3 / 1e6 / gather,
                        "hits": int((d_tags != 0).sum())}
    del km, d_tags
    if fused:
        e.reserve_barcodes(t.n_barcodes)
        ms = []
        for i in range(4):
            e.reset_counts()
            e.sync()
            e.timer_start()
            e.submit_batch_device(d_bases.data_ptr(), n * L, d_off.data_ptr(), d_bc.data_ptr(), n)
            ms.append(e.timer_stop())
        tf = min(ms[1:])
        st = e.stats()
        out["fused_classify"] = {"ms": tf, "lookups": st["lookups"], "lookups_per_s": st["lookups"] / tf * 1e3,
                                 "algorithmic_GBps": st["lookups"] * 32 / tf / 1e6, "filter_pass_frac": st["filter_pass"] / max(1, st["lookups"])}
    e.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--sweep", action="store_true")
    args = ap.parse_args()
    import torch
    from hast_b200 import synth
    from hast_b200.capi import Engine
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    res = {"peak_hbm_copy_GBps": peaks["hbm_gbs"], "cases": []}
    res["cases"].append(dict(measure(21, args.reads, torch, synth, Engine, peaks), name="cfg2 table (128 MiB)"))
    res["cases"].append(dict(measure(21, args.reads, torch, synth, Engine, peaks, decoys=26_800_000), name="cfg3t table (1 GiB, human-scale lists)"))
    if args.sweep:
        for k in (17, 25, 31):
            res["cases"].append(dict(measure(k, args.reads, torch, synth, Engine, peaks), name=f"k sweep k={k}"))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
