#!/bin/bash
# One `ncu --set full` capture of a classify_kernel launch on the cfg2 table (128 MiB) and one on the cfg3 table
# (3.1 Gbp trio, 1 GiB), both through bench.py; the DRAM bytes of each go into profiles/traffic.json, which
# bench.py reads for roofline.traffic / frac_dram.      usage: capture_traffic.sh TAG
O=gpurun_out; T=${1:-r02}; mkdir -p $O
# cfg2: warm-up 3 x 10 launches, capture launch 32
ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^classify_kernel -s 32 -c 1 -f \
    -o $O/${T}_cfg2 python bench.py --no-cpu-baseline --no-e2e --no-cfg3 --steps 1 --warmup 3 > $O/${T}_cfg2_ncu.log 2>&1
# cfg3: the cfg2 leg launches 50 times (3 warm-up + 1 timed + 1 kernel-only, 10 launches each); 16 M pairs = 8 launches per cfg3 step
ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^classify_kernel -s 62 -c 1 -f \
    -o $O/${T}_cfg3 python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 3 --cfg3-pairs 16000000 --cfg3-steps 1 > $O/${T}_cfg3_ncu.log 2>&1
for W in cfg2 cfg3; do
  ncu -i $O/${T}_$W.ncu-rep --page raw --csv > $O/${T}_${W}_raw.csv 2>/dev/null
  ncu -i $O/${T}_$W.ncu-rep --page details > $O/${T}_${W}_details.txt 2>/dev/null
  echo "== $W"; python profiles/tools/ncu_raw.py $O/${T}_${W}_raw.csv
done
python - <<P
import csv, json
out = {"capture": "$T"}
for w, key in (("cfg2", "fused_kernel_dram_bytes_per_launch"), ("cfg3", "fused_kernel_dram_bytes_per_launch_cfg3")):
    rows = list(csv.reader(open("$O/${T}_%s_raw.csv" % w)))
    hdr, units, data = rows[0], rows[1], rows[2]
    def val(name):
        i = hdr.index(name)
        v = float(data[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out[key] = int(rd + wr)
    out[key + "_read"] = int(rd)
    out[key + "_write"] = int(wr)
    out["kernel_" + w] = data[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None
    out["l2_hit_pct_" + w] = float(data[hdr.index("lts__t_sector_hit_rate.pct")])
    out["duration_ns_under_ncu_" + w] = val("gpu__time_duration.sum")
out["source"] = "ncu --set full --clock-control none, one launch of classify_kernel over 4,000,000 reads each: dram__bytes_read.sum + dram__bytes_write.sum (profiles/tools/capture_traffic.sh)"
json.dump(out, open("$O/${T}_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
P
