#!/bin/bash
# per-kernel device times (ncu, serialised) of one rasterizer fwd+bwd at config-5 size
mkdir -p gpurun_out
cat > /tmp/c5.py <<'PY'
import math, os, sys
import torch
ROOT = os.getcwd()
for p in (ROOT, os.path.join(ROOT, "sparse-view-3dgs-pack_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
from lgdwt_b200 import scenes
Wd, Hd = 1920, 1080
fovy = 2 * math.atan(math.tan(0.5) * Hd / Wd)
sc = scenes.trained_like_scene(6_000_000, seed=5, sigma_xyz=1.2, clip=3.0, log_scale_mean=math.log(0.006))
cam = scenes.look_at_camera(Wd, Hd, 1.0, fovy, (0.0, 0.0, -5.0))
t, c = helpers.scene_to_torch(sc), helpers.cam_to_torch(cam)
bg = torch.zeros(3, device="cuda")
dL = torch.randn((3, Hd, Wd), device="cuda")
for _ in range(2):
    f = helpers.run_ours(t, c, cam, bg, want_state=False)
    helpers.backward_ours(t, c, cam, bg, f, dL, None)
torch.cuda.synchronize()
print("done")
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'preprocess|scatter|tile_|blend' --csv --log-file gpurun_out/cfg5_launches.csv python /tmp/c5.py > /tmp/c5.log 2>&1
tail -2 /tmp/c5.log
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/cfg5_launches.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[h]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit'); ii=hdr.index('ID')
recs={}
for r in rows[h+1:]:
    if len(r)<=vi: continue
    recs.setdefault(r[ii],{'k':r[ki].split('(')[0][-40:]})[r[mi]]=(r[vi],r[ui])
ids=sorted(recs,key=int)
for i in ids[len(ids)//2:]:
    d=recs[i]; print("%-42s"%d['k'], {k:v for k,v in d.items() if k!='k'})
PY
