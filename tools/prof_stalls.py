"""dev helper: aggregate ncu warp-stall samples by source line (needs gpurun_out/sass.csv + /tmp/sass/all.sass)."""
import re,csv,collections,sys
kern=sys.argv[1] if len(sys.argv)>1 else 'lift_small'
lines=open('/tmp/sass/all.sass').read().split('\n')
cur_fn=None; cur_line=None; addr2line={}
for ln in lines:
    m=re.match(r'\s*\.section\s+\.text\.(\S+?),',ln)
    if m: cur_fn=m.group(1); continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: cur_line=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',ln)
    if m and cur_fn and kern in cur_fn: addr2line[int(m.group(1),16)]=(cur_line,m.group(2))
rows=list(csv.reader(open('/root/repo/gpurun_out/sass.csv')))
hdr=None; samp=collections.Counter(); inst=collections.Counter(); k=0
stall_cols=None; bystall=collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0]=='Kernel Name': k+=1; continue
    if r and r[0]=='Address':
        hdr=r; isamp=hdr.index('Warp Stall Sampling (All Samples)'); ie=hdr.index('Instructions Executed')
        stall_cols=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if k!=1 or hdr is None or len(r)<=ie: continue
    try: a=int(r[0],16) if r[0].startswith('0x') else int(r[0])
    except: continue
    samp[a]+=int(float(r[isamp] or 0)); inst[a]+=int(float(r[ie] or 0))
    for i,h in stall_cols:
        if i<len(r) and r[i]: bystall[a][h]+=int(float(r[i]))
base=min(samp); tot=sum(samp.values())
byline=collections.Counter(); byline_inst=collections.Counter(); byline_st=collections.defaultdict(collections.Counter)
for a,n in samp.items():
    info=addr2line.get(a-base); key=info[0] if info else None
    byline[key]+=n; byline_inst[key]+=inst[a]
    for h,c in bystall[a].items(): byline_st[key][h]+=c
import glob, os
SRC_FILES = sorted(os.path.basename(f) for f in glob.glob('/root/repo/3d-localisation-and-mapping_b200/csrc/*.cu*'))
src={}
for f in SRC_FILES:
    src[f]=open('/root/repo/3d-localisation-and-mapping_b200/csrc/'+f).read().split('\n')
print('total samples',tot)
tots=collections.Counter()
for key in byline_st:
    for h,c in byline_st[key].items(): tots[h]+=c
print('by reason', [(h,round(100*c/tot,1)) for h,c in tots.most_common(10)])
for key,n in byline.most_common(40):
    if not key: continue
    f,l=key
    text=src.get(f,[''])[l-1].strip()[:60] if f in src and l-1<len(src[f]) else ''
    st=' '.join(f"{h[6:]}:{100*c/n:.0f}" for h,c in byline_st[key].most_common(3))
    print(f"{100*n/tot:5.1f}% inst={100*byline_inst[key]/sum(inst.values()):4.1f}% {f[:12]:12s}{l:5d} {text:60s} | {st}")
