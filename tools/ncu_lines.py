"""Dev tool: per-source-line instruction counts, stall samples and dominant stall reason of one
kernel in an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [n_envs] [lib.so]"""
import collections, csv, os, re, subprocess, sys

rep, kre = sys.argv[1], sys.argv[2]
n_envs = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[4] if len(sys.argv) > 4 else os.path.join(root, "isaacgymloco_b200", "libhimloco_b200.so")
os.makedirs("/tmp/cub", exist_ok=True)
subprocess.run(f"cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all {so} > /dev/null 2>&1", shell=True)
cub = [f for f in os.listdir("/tmp/cub") if f.startswith("hl_env_kernels.sm")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", f"/tmp/cub/{cub}"], capture_output=True, text=True).stdout.split("\n")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append((r[1], cur))
        continue
    if cur is not None:
        cur.append(r)
kname, b = blocks[0]
hdr = b[0]
data = [r for r in b[1:] if len(r) == len(hdr)]
starts = [i for i, l in enumerate(dis) if l.startswith(".text.")]
best = None
for si, s in enumerate(starts):
    e = starts[si + 1] if si + 1 < len(starts) else len(dis)
    if not re.search(kre, dis[s]):
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
    sys.exit(f"could not match disassembly for {kname} ({len(data)} insts)")
ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
I = lambda x: int(x) if x.strip().isdigit() else 0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
byline, samp, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
op = collections.Counter()
for (fl, t), r in zip(best, data):
    byline[fl] += I(r[ie])
    samp[fl] += I(r[ns])
    for s in stalls:
        st[fl][s[6:]] += I(r[hdr.index(s)])
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", t)
    if m:
        op[m.group(2).split(".")[0]] += I(r[ie])
tot, ts = sum(byline.values()), sum(samp.values())
agg = collections.Counter()
for fl in st:
    agg.update(st[fl])
tst = sum(agg.values()) or 1
print(f"kernel: {kname[:100]}")
print(f"sass instructions: {len(data)}; warp-instructions executed: {tot} = {tot / n_envs:.1f} per env")
print("stall reasons: " + ", ".join(f"{k} {100 * v / tst:.1f}%" for k, v in agg.most_common(9)))
print("opcodes/env: " + ", ".join(f"{k} {v / n_envs:.0f}" for k, v in op.most_common(18)))
src = {}
for f in os.listdir(os.path.join(root, "isaacgymloco_b200", "csrc")):
    src[f] = open(os.path.join(root, "isaacgymloco_b200", "csrc", f), errors="ignore").read().split("\n")
print("hot source lines by stall samples (samples %, warp-inst/env, top stalls):")
for fl, nn in samp.most_common(int(os.environ.get("TOPN", 45))):
    s = src.get(fl[0], [""] * 100000)[fl[1] - 1].strip()[:70] if fl and fl[0] in src else ""
    top = ", ".join(f"{k} {100 * v / max(sum(st[fl].values()), 1):.0f}%" for k, v in st[fl].most_common(2))
    print(f"  {100 * nn / ts:5.1f}% {byline[fl] / n_envs:7.1f}  {fl[0] if fl else '?'}:{fl[1] if fl else 0:4d}  [{top}]  {s}")
perfile = collections.Counter()
for fl, nn in byline.items():
    perfile[fl[0] if fl else "?"] += nn
print("per file (warp-inst/env): " + ", ".join(f"{k} {v / n_envs:.0f}" for k, v in perfile.most_common()))
