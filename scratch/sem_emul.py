import numpy as np, oracle
f32=np.float32
BASE=f32(0.10536051565782628)
def e(x): return f32(np.exp(np.float64(x)))
def l(x): return f32(np.log(np.float64(x)))
def lae(a,b):
    if a==-np.inf: return b
    if b==-np.inf: return a
    m=max(a,b); return f32(m+l(f32(e(f32(a-m))+e(f32(b-m)))))
rng = np.random.default_rng(11)
ref = oracle.RefSemanticGrid(0.05, "probabilistic"); ref.set_depth_threshold(1.5); ref.set_depth_decay_rate(0.8)
thr, rate = f32(1.5), f32(0.8)
inv = f32(1.0)/f32(0.05)
vox={}
variants = [dict(f64=True, u8=False, inst=True, depth=True), dict(f64=False, u8=True, inst=True, depth=False),
            dict(f64=True, u8=False, inst=False, depth=True), dict(f64=False, u8=False, inst=False, depth=False)]
for var in variants:
    n=30000
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    pts = dirs * (0.5 + 0.01 * rng.normal(size=(n, 1))) + [0.05, -0.1, 0.02]
    pts = pts.astype(np.float64 if var["f64"] else np.float32)
    cols_u8 = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    cols_f = (cols_u8.astype(np.float32) * (np.float32(1.0) / np.float32(255.0))) if var["u8"] else rng.random((n, 3)).astype(np.float32)
    side = (pts[:, 0] > 0).astype(np.int32)
    flip, noise = rng.random(n) < 0.2, rng.integers(-1, 1, n)
    cls = np.where(flip, noise, 1 + side).astype(np.int32)
    ins = np.where(flip, noise, 10 + side).astype(np.int32)
    dep = rng.uniform(0.5, 4.0, n).astype(np.float32)
    ref.integrate(pts, cols_f, cls, ins if var["inst"] else None, dep if var["depth"] else None)
    if var["f64"]: key = np.floor(pts * np.float64(inv)).astype(np.int64)
    else: key = np.floor(pts * inv).astype(np.int64)
    for i in range(n):
        k=tuple(key[i]); st=vox.get(k)
        if st is None: st=vox[k]=dict(count=0,labs={},ml=(-1,-1),mlp=f32(-np.inf))
        oo = int(ins[i]) if var["inst"] else 0; oc=int(cls[i])
        w=BASE
        if var["depth"] and not (dep[i] <= thr):
            w = f32(e(f32(f32(-(f32(dep[i]-thr)))*rate))*BASE)
        pair=(oo,oc)
        if st["count"]==0:
            st["labs"][pair]=w; st["ml"]=pair; st["mlp"]=w
        elif pair in st["labs"]:
            st["labs"][pair]=f32(st["labs"][pair]+w)
            if pair==st["ml"]: st["mlp"]=st["labs"][pair]
            elif st["labs"][pair]>st["mlp"]: st["mlp"]=st["labs"][pair]; st["ml"]=pair
        else:
            st["labs"][pair]=w
            if w>st["mlp"]: st["mlp"]=w; st["ml"]=pair
        st["count"]+=1
d=ref.dump_blocks(8)
keys=d["keys"]; worst=0; nbad=0; tot=0
for b in range(len(keys)):
    for li in np.where(d["count"][b]>0)[0]:
        lx,ly,lz=li&7,(li>>3)&7,li>>6
        k=(keys[b,0]*8+lx,keys[b,1]*8+ly,keys[b,2]*8+lz)
        st=vox[k]; tot+=1
        assert st["count"]==d["count"][b,li]
        s=f32(-np.inf)
        for p in sorted(st["labs"]): s=lae(s,st["labs"][p])
        c = f32(0) if -1 in st["ml"] else e(f32(st["mlp"]-s))
        rc=d["confidence"][b,li]
        if st["ml"]!=(d["object_id"][b,li],d["class_id"][b,li]): print("label mismatch",k,st["ml"],d["object_id"][b,li],d["class_id"][b,li])
        r=abs(float(c)-float(rc))/max(float(rc),1e-12)
        if abs(float(c)-float(rc))>1e-9+2e-6*float(rc):
            nbad+=1
            if nbad<5: print("conf mismatch",k,c,rc,st["labs"],st["ml"],st["mlp"], list(zip(d["lab_obj"][b,li],d["lab_cls"][b,li],d["lab_logp"][b,li]))[:6])
        worst=max(worst,r)
print("voxels",tot,"worst rel",worst,"bad",nbad)
nd=0
for b in range(len(keys)):
    for li in np.where(d["count"][b]>0)[0]:
        lx,ly,lz=li&7,(li>>3)&7,li>>6
        k=(keys[b,0]*8+lx,keys[b,1]*8+ly,keys[b,2]*8+lz)
        st=vox[k]
        mine=[st["labs"][p] for p in sorted(st["labs"])]
        refl=list(d["lab_logp"][b,li][:len(mine)])
        if any(float(x)!=float(y) for x,y in zip(mine,refl)): nd+=1
print("voxels with logp differing (correctly-rounded exp vs glibc):", nd)
