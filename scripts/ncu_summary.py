"""Summarise an .ncu-rep: key metrics per kernel and (optionally) the top stall lines.
usage: python scripts/ncu_summary.py rep.ncu-rep [--top N] [--kernel REGEX] [--json WORKLOAD:STORAGE --commit SHA]
--json merges {name, time_ms, dram_read, dram_write, l1_lsu_wavefront_pct, regs} of every kernel (first launch of each
name) into profiles/ncu_kernels.json under that key: bench.py reads `roofline.traffic` from there."""
import csv, subprocess, sys, io, re, collections, json, os
rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
kre = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else None
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
names = []
if "--json" in sys.argv:
    key = sys.argv[sys.argv.index("--json") + 1]
    commit = sys.argv[sys.argv.index("--commit") + 1] if "--commit" in sys.argv else None
    def num(r, k, scale=1.0):
        if k not in hdr or not r[hdr.index(k)]:
            return None
        v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
        return v * mult * scale
    seen, ks = set(), []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if name in seen:
            continue
        seen.add(name)
        ks.append({"name": name, "time_ms": num(r, "gpu__time_duration.sum"), "dram_read": num(r, "dram__bytes_read.sum"),
                   "dram_write": num(r, "dram__bytes_write.sum"),
                   "l1_lsu_wavefront_pct": num(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                   "lts_hit_pct": num(r, "lts__t_sector_hit_rate.pct"), "regs": num(r, "launch__registers_per_thread"),
                   "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                   "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_kernels.json")
    db = json.load(open(out)) if os.path.exists(out) else {}
    db[key] = {"file": os.path.basename(rep), "commit": commit, "kernels": ks}
    json.dump(db, open(out, "w"), indent=1)
    print("wrote", out, key, len(ks), "kernels")
    sys.exit(0)
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    names.append(name)
    if kre and not re.search(kre, name):
        continue
    print("##", name[:100])
    for k in KEYS:
        if k in hdr:
            print(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: float(r[i]) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]}
    print("  stalls/issue:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
if top:
    for idx, name in enumerate(names):
        if kre and not re.search(kre, name):
            continue
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{re.escape(re.split(r'[<(]', name)[0].split()[-1])}:{1}"],
                             capture_output=True, text=True).stdout
        rs = list(csv.reader(io.StringIO(src)))
        if len(rs) < 3:
            continue
        h = rs[1]
        ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        data = []
        for k, r in enumerate(rs[2:]):
            if len(r) > isamp and r[isamp].isdigit():
                best = max(stall_cols, key=lambda i: int(r[i]) if r[i].isdigit() else 0)
                data.append((int(r[isamp]), k, r[ia].strip()[:70], h[best], int(r[iex])))
        tot = sum(d[0] for d in data) or 1
        print("## top stall lines of", rs[0][1][:80], "total samples", tot)
        for d in sorted(data, reverse=True)[:top]:
            print(f"  {100 * d[0] / tot:5.1f}%  line {d[1]:5d}  {d[2]:70s} {d[3]} exec={d[4]}")
