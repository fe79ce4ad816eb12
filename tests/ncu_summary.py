"""turn an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into the text summary kept under profiles/ and
   the per-kernel dram traffic table profiles/ncu_traffic.json that bench.py reads for roofline.traffic"""
import csv, json, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum"]
def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
traffic = {}
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("==== %s" % rep)
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print("-----\n   Kernel Name = %s" % name[:150])
        for w in WANT:
            if w in col:
                print("   %s = %s %s" % (w, r[col[w]], units[col[w]]))
        short = name.split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        traffic.setdefault(short, rd + wr)
json.dump(traffic, open("profiles/ncu_traffic.json", "w"), indent=1, sort_keys=True)
