import os, re, sys, tempfile, pathlib, numpy as np
os.environ["EMD_SNAP_ENERGY"]="1"
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import examinimd_b200 as emd
REPO=pathlib.Path("/root/repo"); SNAP_DIR=REPO/"input"/"snap"
td=pathlib.Path(tempfile.mkdtemp())
txt=(SNAP_DIR/"in.snap.W").read_text()
txt=re.sub(r"region\s+box block.*","region\t\tbox block 0 4 0 4 0 4",txt)
(td/"in.deck").write_text(txt)
for f in SNAP_DIR.glob("*.snap*"): (td/f.name).write_bytes(f.read_bytes())
app=emd.App(["-il",str(td/"in.deck"),"--neigh-type","CSR","--comm-type","SERIAL"])
n=app.get("N_local"); dt=0.001
pes=[];kes=[];fv=[]
for s in range(70):
    T,pe,ke=app.thermo(); cur=app.download()
    pes.append(pe*n);kes.append(ke*n);fv.append((cur["f"]*cur["v"]).sum()); app.advance(1)
pes=np.array(pes);kes=np.array(kes);fv=np.array(fv)
dpe=(pes[2:]-pes[:-2])/(2*dt)
np.set_printoptions(linewidth=200,precision=5)
print("rel err", (dpe+fv[1:-1])/np.abs(fv).max())
print("etot-e0", (pes+kes)-(pes+kes)[0]); print("ke mean",kes.mean())
