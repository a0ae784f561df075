import sys,re,csv,collections,subprocess
rep=sys.argv[1]; n_envs=65536
subprocess.run("mkdir -p /tmp/cub && cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all /root/repo/isaacgymloco_b200/libhimloco_b200.so > /dev/null 2>&1", shell=True)
dis=subprocess.run(["nvdisasm","-g","-c","/tmp/cub/hl_env_kernels.sm_100a.cubin"],capture_output=True,text=True).stdout.split("\n")
txt=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
blocks=[];cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur=[];blocks.append(cur);continue
    if cur is not None: cur.append(r)
b=blocks[0];hdr=b[0];data=[r for r in b[1:] if len(r)==len(hdr)]
starts=[i for i,l in enumerate(dis) if l.startswith(".text.")]
for si,s in enumerate(starts):
    e=starts[si+1] if si+1<len(starts) else len(dis)
    if 'fused' not in dis[s]: continue
    insts=[];curl=None
    for l in dis[s:e]:
        m=re.search(r'//## File "([^"]+)", line (\d+)',l)
        if m: curl=(m.group(1).split('/')[-1],int(m.group(2)));continue
        m=re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);",l)
        if m: insts.append((curl,m.group(2)))
    if len(insts)==len(data): break
I=lambda x:int(x) if x.strip().isdigit() else 0
ie=hdr.index("Instructions Executed")
# group consecutive instructions by per-env execution ratio bucket
seq=[(I(r[ie])/n_envs, fl, t) for (fl,t),r in zip(insts,data)]
# segments: runs where ratio within 5%
segs=[];cur=None
for ratio,fl,t in seq:
    key=round(ratio,2)
    if cur and abs(cur[0]-ratio)<0.03*max(cur[0],0.05): cur[1]+=1; cur[2]+=ratio; cur[3].append((fl,t))
    else:
        cur=[ratio,1,ratio,[(fl,t)]]; segs.append(cur)
print("segments (exec/env ratio, #sass, total inst/env, first line .. last line):")
for r,n,tot,lst in segs:
    if tot<3: continue
    lines=[x[0][1] for x in lst if x[0] and x[0][0]=='hl_env_kernels.cu']
    ops=collections.Counter(re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)",x[1]).group(2) for x in lst)
    print(f"  ratio {r:6.3f} n={n:4d} tot={tot:7.1f}  lines {min(lines) if lines else '?'}..{max(lines) if lines else '?'}  top: {', '.join(f'{k}{v}' for k,v in ops.most_common(8))}")
