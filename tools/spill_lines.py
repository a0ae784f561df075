"""Dev tool: where the local-memory (spill) loads/stores of a kernel are, by source line.
usage: python tools/spill_lines.py <lib.so> <function-name regex>"""
import collections, os, re, subprocess, sys
so, fre = sys.argv[1], sys.argv[2]
os.makedirs("/tmp/cub2", exist_ok=True)
subprocess.run(f"cd /tmp/cub2 && rm -f *.cubin && cuobjdump -xelf all {os.path.abspath(so)} > /dev/null 2>&1", shell=True)
cub = [f for f in os.listdir("/tmp/cub2") if f.startswith("hl_env_kernels.sm")][0]
lines = subprocess.run(["nvdisasm", "-g", "-c", f"/tmp/cub2/{cub}"], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and re.search(fre, l))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.")), len(lines))
cur, cnt, tot = None, collections.Counter(), 0
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        tot += 1
        if re.search(r"\b(LDL|STL)", m.group(2)):
            cnt[cur] += 1
print("instructions:", tot, " spill ld/st:", sum(cnt.values()))
byfile = collections.Counter()
for (f, ln), v in cnt.items():
    byfile[(f, ln // 50 * 50)] += v
for k, v in sorted(byfile.items()):
    print(f"  {k[0]}:{k[1]}-{k[1]+49}: {v}")
