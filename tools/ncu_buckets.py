import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur=None
b={}
def bucket(f,l):
    l=int(l)
    if f=='trace_device.cuh':
        if l<=46: return 'math'
        if l<=70: return 'philox'
        if l<=84: return 'ld'
        if l<=100: return 'toLocal'
        if l<=186: return 'intersect'
        if 270<=l<=283: return 'trav_setup'
        if 284<=l<=316: return 'leaf'
        if 317<=l<=356: return 'nodeloop'
        if 357<=l<=372: return 'leafphase'
        if 373<=l<=440: return 'surface'
        if l<=472: return 'tex'
        return 'material'
    if f=='trace_kernels.cu':
        if l<=123: return 'k_fetch'
        if l<=142: return 'k_gen'
        if l<=148: return 'k_travcall'
        if l<=162: return 'k_miss'
        if l<=200: return 'k_shade'
        return 'k_tail'
    return f
T=[0,0,0]
for r in rows:
    if len(r)>=2 and r[0]=="File Path":
        cur=r[1].split("/")[-1]; continue
    if len(r)>10 and r[0] not in ("","Line No") and r[2]=="-":
        try: inst,thr,samp=int(r[7]),int(r[8]),int(r[6])
        except ValueError: continue
        k=bucket(cur,r[0])
        x=b.setdefault(k,[0,0,0]); x[0]+=inst;x[1]+=thr;x[2]+=samp
        T[0]+=inst;T[1]+=thr;T[2]+=samp
for k,(i,t,s) in sorted(b.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:12s} inst {100*i/T[0]:5.1f}%  thr-inst {100*t/T[1]:5.1f}%  samp {100*s/T[2]:5.1f}%  avg thr {t/max(i,1):5.1f}  lost-slots {100*(32*i-t)/(32*T[0]):5.1f}%")
print(T, T[1]/T[0])
