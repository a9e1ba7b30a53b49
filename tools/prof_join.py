import re,csv,collections,sys
kern=sys.argv[1]; rep_csv=sys.argv[2]; slots=float(sys.argv[3]) if len(sys.argv)>3 else 1.0
lines=open('/tmp/sass/all.sass').read().split('\n')
cur_fn=None; cur_line=None; addr2line={}
for ln in lines:
    m=re.match(r'\s*\.section\s+\.text\.(\S+?),',ln)
    if m: cur_fn=m.group(1); continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: cur_line=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',ln)
    if m and cur_fn and kern in cur_fn: addr2line[int(m.group(1),16)]=(cur_line,m.group(2))
rows=list(csv.reader(open(rep_csv)))
hdr=None; agg=collections.Counter(); k=0
for r in rows:
    if r and r[0]=='Kernel Name': k+=1; continue
    if r and r[0]=='Address': hdr=r; ie=hdr.index('Instructions Executed'); continue
    if k!=1 or hdr is None or len(r)<=ie: continue
    try: a=int(r[0],16) if r[0].startswith('0x') else int(r[0])
    except: continue
    agg[a]+=int(float(r[ie] or 0))
base=min(agg); tot=sum(agg.values())
byline=collections.Counter(); byop=collections.defaultdict(collections.Counter)
for a,n in agg.items():
    info=addr2line.get(a-base)
    key=info[0] if info else None
    byline[key]+=n
    if info: byop[key][info[1].split()[0] if not info[1].startswith('@') else info[1].split()[1]]+=n
import glob, os
SRC_FILES = sorted(os.path.basename(f) for f in glob.glob('/root/repo/3d-localisation-and-mapping_b200/csrc/*.cu*'))
src={}
for f in SRC_FILES:
    src[f]=open('/root/repo/3d-localisation-and-mapping_b200/csrc/'+f).read().split('\n')
print('total',tot)
for f in SRC_FILES:
    for (ff,l),n in sorted((k,v) for k,v in byline.items() if k and k[0]==f):
        if n/tot<0.002: continue
        ops=' '.join(f"{o}:{c/slots:.1f}" for o,c in byop[(ff,l)].most_common(6))
        print(f"{f[:12]:12s}{l:5d} {100*n/tot:5.1f}% {n/slots:7.2f}  {src[f][l-1].strip()[:70]:70s} | {ops}")
other=sum(n for k,n in byline.items() if not k or k[0] not in src)
print('other files',other/tot)
for k,n in byline.most_common():
    if k and k[0] not in src and n/tot>0.005: print(k,n/slots, dict(byop[k].most_common(4)))
