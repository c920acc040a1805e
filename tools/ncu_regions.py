import csv,re,sys
sys.path.insert(0,'/root/repo/tools')
from ncu_by_line import parse_disasm
def load(path): return open(path).read().split('\n')
srcI=load('/root/repo/path_trace_golang_b200/csrc/integrator.cu'); srcW=load('/root/repo/path_trace_golang_b200/csrc/wavefront.cuh')
def find(src,pat,start=0): return next(i for i,l in enumerate(src) if i>=start and pat in l)+1
mi=[("rng",find(srcI,"struct Rng")-8),("vec helpers",find(srcI,"struct F3")),("in_unit_sphere",find(srcI,"F3 in_unit_sphere")),("cosine_direction",find(srcI,"F3 cosine_direction")-1),
("make_ray",find(srcI,"struct RayK")-3),("hit box",find(srcI,"bool hit_box")-7),("hit sphere",find(srcI,"bool hit_sphere")-1),("hit plane",find(srcI,"bool hit_plane")-1),("obj load/hit_any",find(srcI,"float4 obj_lo")-1),
("surface",find(srcI,"void surface")-1),("sky",find(srcI,"F3 sky_color")),("to_u8",find(srcI,"uint8_t to_u8")-1),("megakernel",find(srcI,"integrate_kernel(const __grid_constant__")-2)]
mw=[("wf prologue",1),("wf regen",find(srcW,"void path_regen")-1),("wf shade: load+surface",find(srcW,"void path_shade")-1),("wf shade: common",find(srcW,"draws consumed by this bounce")),
("wf shade: diffuse",find(srcW,"} else if (c == CL_DIFFUSE)")),("wf shade: dielectric",find(srcW,"} else if (c == CL_DIEL)")),("wf exit search",find(srcW,"if (front) {")),
("wf RR/update",find(srcW,"bool done = !ok;")),("wf term",find(srcW,"} else if (c == CL_TERM")),("wf kernel prologue",find(srcW,"integrate_wf_kernel(const __grid_constant__")-2),("wf scan",find(srcW,"SCAN (thread")),("wf sort",find(srcW,"SORT (stable")),("wf shade loop+barrier",find(srcW,"SHADE (one class per")),("wf epilogue",find(srcW,"if (STATS) {",find(srcW,"SHADE (one class per")))]
def region(f,line):
    marks = mi if f=="integrator.cu" else mw if f=="wavefront.cuh" else None
    if marks is None: return "other:"+f
    name="other"
    for n,l in sorted(marks,key=lambda x:x[1]):
        if line>=l: name=n
    return name
dmap=parse_disasm(sys.argv[2],sys.argv[4])
rows=list(csv.reader(open(sys.argv[1])))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="Address"); hdr=rows[hi]; col={n:i for i,n in enumerate(hdr)}
data=rows[hi+1:]; base=int(data[0][0],16)
nsamp=float(sys.argv[3])
agg={}; tot=0
for r in data:
    off=int(r[0],16)-base; inst=int(r[col["Instructions Executed"]] or 0); tin=int(r[col["Thread Instructions Executed"]] or 0); smp=int(r[col["# Samples"]] or 0)
    key=dmap.get(off,(None,""))[0]
    name=region(key[0],key[1]) if key else "other"
    a=agg.setdefault(name,[0,0,0]); a[0]+=inst; a[1]+=tin; a[2]+=smp; tot+=inst
ts=sum(a[2] for a in agg.values())
print(f"total warp-inst/sample {tot/nsamp:.1f}")
for n,a in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    if a[0]: print(f"{n:32s} {100*a[0]/tot:6.2f}%  active {a[1]/a[0]:5.1f}  warp-inst/sample {a[0]/nsamp:6.1f}  samples {100*a[2]/ts:5.1f}%")
