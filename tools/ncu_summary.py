"""Dev tool: summarise gpurun_out artefacts (launch list, full ncu report joined with nvdisasm line info)."""
import collections
import csv
import json
import os
import pickle
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0][:48]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = []
    for k, a in agg.items():
        out.append(f"{k:50s} n={a[0]:4d} avg={a[1]/a[0]:9.1f} us share={100*a[1]/tot:5.1f}%")
    return "\n".join(out)


def raw_metrics(rep, names):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for w in names:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"{w:68s} {data[0][i]:>16s} {units[i]}")
    return "\n".join(out)


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
           "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__inst_executed.sum", "launch__waves_per_multiprocessor"]


def sass_join(rep, so, kernel_re, n_envs):
    os.makedirs("/tmp/cub", exist_ok=True)
    subprocess.run(f"cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all {so} > /dev/null 2>&1", shell=True)
    cub = [f for f in os.listdir("/tmp/cub") if f.startswith("hl_env_kernels.sm")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", f"/tmp/cub/{cub}"], capture_output=True, text=True).stdout.split("\n")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    kname = [r[1] for r in rows if r and r[0] == "Kernel Name"][0]
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
            continue
        if cur is not None:
            cur.append(r)
    b = blocks[0]
    hdr = b[0]
    data = [r for r in b[1:] if len(r) == len(hdr)]
    # find the matching function in the disassembly by instruction count
    starts = [i for i, l in enumerate(dis) if l.startswith(".text.")]
    best = None
    for si, s in enumerate(starts):
        e = starts[si + 1] if si + 1 < len(starts) else len(dis)
        if not re.search(kernel_re, dis[s]):
            continue
        insts, curl = [], None
        for l in dis[s:e]:
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                curl = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
            if m:
                insts.append((curl, m.group(2)))
        if len(insts) == len(data):
            best = insts
            break
    if best is None:
        return f"could not match disassembly for {kname} ({len(data)} insts)"
    ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
    I = lambda x: int(x) if x.strip().isdigit() else 0
    byline, samp, op = collections.Counter(), collections.Counter(), collections.Counter()
    for (fl, txt_), r in zip(best, data):
        byline[fl] += I(r[ie])
        samp[fl] += I(r[ns])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", txt_)
        if m:
            op[m.group(2).split(".")[0]] += I(r[ie])
    tot, ts = sum(byline.values()), sum(samp.values())
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(I(r[hdr.index(s)]) for r in data) for s in stalls}
    tst = sum(agg.values()) or 1
    out = [f"kernel: {kname[:90]}", f"sass instructions: {len(data)}; warp-instructions executed: {tot} = {tot/n_envs:.1f} per env",
           "stall reasons: " + ", ".join(f"{s[6:]} {100*v/tst:.1f}%" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]),
           "opcodes/env: " + ", ".join(f"{k} {v/n_envs:.0f}" for k, v in op.most_common(16)), "hot source lines (warp-inst/env, stall-sample share):"]
    src = {}
    for f in ("hl_env_kernels.cu", "hl_math.cuh", "hl_common.cuh", "hl_fused_kernel.inc"):
        src[f] = open(os.path.join(os.path.dirname(so), "csrc", f)).read().split("\n")
    for fl, nn in byline.most_common(int(os.environ.get("TOPN", 40))):
        s = src.get(fl[0], [""] * 10000)[fl[1] - 1].strip()[:80] if fl and fl[0] in src else ""
        out.append(f"  {nn/n_envs:7.1f} {100*samp[fl]/ts:5.1f}%  {fl[0] if fl else '?'}:{fl[1] if fl else 0:4d}  {s}")
    # per-file totals and, for the fused kernel body, per-phase totals (markers in the source)
    perfile = collections.Counter()
    for fl, nn in byline.items():
        perfile[fl[0] if fl else "?"] += nn
    out.append("per file (warp-inst/env): " + ", ".join(f"{k} {v/n_envs:.0f}" for k, v in perfile.most_common()))
    inc = src.get("hl_fused_kernel.inc")
    if inc:
        marks = [(i + 1, l.strip()[:60]) for i, l in enumerate(inc) if "// ----------------" in l]
        marks = [(0, "prologue")] + marks + [(10 ** 9, "")]
        for (a, name), (b_, _) in zip(marks[:-1], marks[1:]):
            tot_r = sum(nn for fl, nn in byline.items() if fl and fl[0] == "hl_fused_kernel.inc" and a <= fl[1] < b_)
            sm_r = sum(nn for fl, nn in samp.items() if fl and fl[0] == "hl_fused_kernel.inc" and a <= fl[1] < b_)
            out.append(f"  .inc lines {a:4d}-: {tot_r/n_envs:7.1f} inst/env, {100*sm_r/ts:5.1f}% samples  {name}")
    return "\n".join(out)


if __name__ == "__main__":
    rep = sys.argv[1]
    n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.path.exists(os.path.join(root, "gpurun_out", "launches.csv")):
        print(launches(os.path.join(root, "gpurun_out", "launches.csv")))
    print(raw_metrics(rep, METRICS))
    print(sass_join(rep, os.path.join(root, "isaacgymloco_b200", "libhimloco_b200.so"), "fused", n_envs))
