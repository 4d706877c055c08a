"""Fold an `ncu --page source --csv --print-source sass,cuda` dump by source function: instructions executed, stall samples and\naverage active lanes per function and kernel.  usage: fold_by_function.py dump.csv"""
import csv,sys,re,collections,bisect
# function line ranges in a source file (crude: top-level lines starting with CATAN_FN / template / __global__)
def func_ranges(path):
    names=[];starts=[]
    for n,l in enumerate(open(path),1):
        m=re.match(r'^(?:template.*>\s*)?(?:CATAN_FN_NOINLINE|CATAN_FN|CATAN_MFN|__global__|__device__|static inline)\b.*?(\w+)\s*\(',l)
        if m and not l.startswith(' '):
            names.append(m.group(1)); starts.append(n)
            continue
        m=re.match(r'^struct\s+(?:alignas\(\d+\)\s+)?(\w+)',l)          # member functions count for their struct
        if m:
            names.append(m.group(1)); starts.append(n)
    return starts,names
cache={}
def fn(path,line):
    if path not in cache:
        try: cache[path]=func_ranges(path)
        except Exception: cache[path]=([],[])
    s,nm=cache[path]
    i=bisect.bisect_right(s,line)-1
    return nm[i] if i>=0 else '?'
rows=csv.reader(open(sys.argv[1]))
cur_file=None;cur_kernel=None;hdr=None
agg=collections.defaultdict(lambda: collections.defaultdict(lambda:[0,0,0]))
mode=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1]; continue
    if r[0]=='Function Name': cur_kernel=r[1]; continue
    if r[0]=='Line No': hdr=r; mode='cuda'; continue
    if r[0]=='Address': hdr=r; mode='sass'; continue
    if mode=='cuda' and hdr:
        try:
            line=int(r[0]); ie=hdr.index('Instructions Executed'); ns=hdr.index('# Samples'); te=hdr.index('Thread Instructions Executed')
            inst=int(r[ie] or 0); samp=int(r[ns] or 0); tinst=int(r[te] or 0)
        except Exception: continue
        f=fn(cur_file,line)+' ['+cur_file.split('/')[-1]+']'
        a=agg[cur_kernel][f]; a[0]+=inst; a[1]+=samp; a[2]+=tinst
for k,v in agg.items():
    ti=sum(a[0] for a in v.values()); ts=sum(a[1] for a in v.values())
    print('==',k,'instr',ti,'samples',ts)
    for f,a in sorted(v.items(),key=lambda x:-x[1][1])[:25]:
        print('  %-50s inst %9d (%4.1f%%)  samples %7d (%4.1f%%)  lanes %.1f'%(f,a[0],100*a[0]/max(ti,1),a[1],100*a[1]/max(ts,1),a[2]/max(a[0],1)))
