"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01        # -> profiles/r01_step_<workload>.md, profiles/r01_ncu_full_<name>.md, profiles/r01_traffic.json
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def short(name):
    return name.split("(")[0].replace("void ", "").replace("icaf::", "").strip()


def step_summary(wl):
    path = os.path.join(GO, f"step_{wl}.csv")
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    c = {h: hdr.index(h) for h in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in data:
        d = launches.setdefault(r[c["ID"]], {"kernel": short(r[c["Kernel Name"]]), "grid": r[c["Grid Size"]].replace(" ", "")})
        v = float(r[c["Metric Value"]].replace(",", ""))
        unit = r[c["Metric Unit"]]
        name = r[c["Metric Name"]]
        if name.startswith("gpu__time_duration"):
            v = v / 1e3 if unit == "ns" else (v if unit in ("us", "usecond") else v * 1e3 if unit == "ms" else v)
            d["us"] = v
        elif name.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["rd" if "read" in name else "wr"] = v * mult
        elif "pipe_tensor" in name:
            d["tensor_pct"] = v
        elif "sm__throughput" in name:
            d["sm_pct"] = v
    L = list(launches.values())
    tot = sum(x["us"] for x in L)
    agg = collections.OrderedDict()
    for x in L:
        a = agg.setdefault(x["kernel"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1; a["us"] += x["us"]; a["rd"] += x.get("rd", 0); a["wr"] += x.get("wr", 0)
    md = [f"# {tag}: one eager step of {wl} under `ncu --metrics gpu__time_duration,dram__bytes_*,tensor pipe` (cold-cache, serialised)",
          "", f"{len(L)} launches, sum of kernel durations {tot:.1f} us. Shares are what matters (ncu flushes caches between kernels).", "",
          "| kernel | launches | total us | share | avg us | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        md.append(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['us'] / a['n']:.1f} | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} |")
    md += ["", "## every launch", "", "| # | kernel | grid | us | DRAM rd MB | DRAM wr MB | tensor pipe % | SM % |", "|---|---|---|---|---|---|---|---|"]
    for i, x in enumerate(L):
        md.append(f"| {i} | `{x['kernel']}` | {x['grid']} | {x['us']:.1f} | {x.get('rd', 0) / 1e6:.2f} | {x.get('wr', 0) / 1e6:.2f} | "
                  f"{x.get('tensor_pct', 0):.1f} | {x.get('sm_pct', 0):.1f} |")
    open(os.path.join(OUT, f"{tag}_step_{wl}.md"), "w").write("\n".join(md) + "\n")
    conv = [x for x in L if "conv_gemm_" in x["kernel"]]
    return {"launches": len(L), "sum_us": tot, "conv_launches": len(conv), "conv_us": sum(x["us"] for x in conv),
            "conv_dram_bytes_per_launch": sum(x.get("rd", 0) + x.get("wr", 0) for x in conv) / max(1, len(conv)),
            "conv_dram_bytes_per_step": sum(x.get("rd", 0) + x.get("wr", 0) for x in conv)}


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]


def full_summary(name):
    rep = os.path.join(GO, name + ".ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    md = [f"# {tag}: `ncu --set full --clock-control none --import-source on` -- {name}", "",
          "Selected raw metrics per captured launch (the .ncu-rep itself is scratch; this table is the tracked evidence).", ""]
    cols = [(h, hdr.index(h)) for h in WANT if h in hdr]
    md.append("| launch | kernel | grid | " + " | ".join(h for h, _ in cols) + " |")
    md.append("|---|---|---|" + "---|" * len(cols))
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    for n, r in enumerate(data):
        md.append(f"| {n} | `{short(r[ki])}` | {r[gi].replace(' ', '')} | " + " | ".join(f"{r[i]} {units[i]}".strip() for _, i in cols) + " |")
    # stall-reason totals from the source page of the first launch
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = [r for r in csv.reader(io.StringIO(src)) if len(r) > 10]
    if srows:
        sh = srows[0]
        body = [r for r in srows[1:] if r[0] != sh[0]]
        sums = {}
        for i, h in enumerate(sh):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    sums[h] = sum(float(r[i]) for r in body if r[i] not in ("", "-"))
                except ValueError:
                    pass
        tot = sum(sums.values()) or 1
        md += ["", "## warp-stall samples (source page, all captured launches of this report)", "", "| reason | samples | share |", "|---|---|---|"]
        for k, v in sorted(sums.items(), key=lambda kv: -kv[1])[:10]:
            md.append(f"| {k} | {v:.0f} | {100 * v / tot:.1f}% |")
        sass = " ".join(r[sh.index("Source")] for r in body if "Source" in sh)
        hits = {m: sass.count(m) for m in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "LDGSTS", "SYNCS", "UCGABAR")}
        md += ["", "SASS mnemonic counts in the captured kernel text: " + ", ".join(f"{k} {v}" for k, v in hits.items())]
    open(os.path.join(OUT, f"{tag}_ncu_full_{name.replace('full_', '')}.md"), "w").write("\n".join(md) + "\n")


os.makedirs(OUT, exist_ok=True)
traffic = {}
for wl in ("yolov5s_b1", "yolov5l_b16"):
    s = step_summary(wl)
    if s:
        traffic[wl] = s
if traffic:
    json.dump(traffic, open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
for name in ("full_conv_l_b16_head", "full_conv_l_b16_pair", "full_conv_l_b16_pair2", "full_conv_s_b1", "full_attn_l_b16", "full_attn_big"):
    full_summary(name)
print(json.dumps(traffic, indent=1))
print(sorted(os.listdir(OUT)))
