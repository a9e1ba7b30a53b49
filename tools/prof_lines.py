"""dev helper: per-source-line instruction / stall-sample shares of one kernel from an ncu report.
usage: prof_lines.py <kernel substring> <sass.csv from ncu --page source --print-source sass> <src.csv from --print-source cuda> [top]"""
import re,csv,collections,sys
kern=sys.argv[1]; sass_csv=sys.argv[2]; src_csv=sys.argv[3]; top=int(sys.argv[4]) if len(sys.argv)>4 else 40
lines=open('/tmp/sass/all.sass').read().split('\n')
cur_fn=None; cur_line=None; addr2line={}
for ln in lines:
    m=re.match(r'\s*\.section\s+\.text\.(\S+?),',ln)
    if m: cur_fn=m.group(1); continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: cur_line=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',ln)
    if m and cur_fn and kern in cur_fn: addr2line[int(m.group(1),16)]=(cur_line,m.group(2))
rows=list(csv.reader(open(sass_csv)))
hdr=None; samp=collections.Counter(); inst=collections.Counter(); bystall=collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0]=='Address':
        hdr=r; isamp=hdr.index('Warp Stall Sampling (All Samples)'); ie=hdr.index('Instructions Executed')
        stall_cols=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]; continue
    if hdr is None or len(r)<=ie: continue
    try: a=int(r[0],16) if r[0].startswith('0x') else int(r[0])
    except: continue
    samp[a]+=int(float(r[isamp] or 0)); inst[a]+=int(float(r[ie] or 0))
    for i,h in stall_cols:
        if i<len(r) and r[i]: bystall[a][h]+=int(float(r[i]))
base=min(samp); tot=sum(samp.values()); toti=sum(inst.values())
byline=collections.Counter(); byline_inst=collections.Counter(); byline_st=collections.defaultdict(collections.Counter)
for a,n in samp.items():
    info=addr2line.get(a-base); key=info[0] if info else None
    byline[key]+=n; byline_inst[key]+=inst[a]
    for h,c in bystall[a].items(): byline_st[key][h]+=c
src={}
cur=None
for r in csv.reader(open(src_csv)):
    if len(r)==2 and r[0]=='File Name': cur=r[1].split('/')[-1]; src[cur]={}
    elif len(r)==2 and r[0].isdigit() and cur: src[cur][int(r[0])]=r[1]
print('total samples',tot,'warp-inst',toti)
tots=collections.Counter()
for a in bystall:
    for h,c in bystall[a].items(): tots[h]+=c
print([(h[6:],round(100*c/tot,1)) for h,c in tots.most_common(8)])
for key,n in sorted(byline.items(), key=lambda kv:-byline_inst[kv[0]])[:top]:
    if not key: continue
    f,l=key
    text=src.get(f,{}).get(l,'').strip()[:78]
    st=' '.join(f"{h[6:]}:{100*c/max(n,1):.0f}" for h,c in byline_st[key].most_common(2))
    print(f"inst {100*byline_inst[key]/toti:5.1f}% samp {100*n/tot:5.1f}% {f[5:17]:12s}{l:5d} {text:78s} | {st}")
