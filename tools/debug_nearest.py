import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests'); sys.path.insert(0,'oracle')
import transform_b200 as t, oracle_np as o
from conftest import to_oracle
ny,nx=300,411
src=np.random.default_rng(3).random((ny,nx)).astype(np.float32)
tr=t.Composition([t.Scale(1.3,dim=2),t.Rotation(30,unit='deg'),t.Offset([-nx/2.0,-ny/2.0])])
for tuning in (0,4):
    got=tr.resample(src,method='nearest',precision='fp32',tuning=tuning)
    want=o.resample(to_oracle(tr),src,method='nearest').astype(np.float32)
    bad=np.argwhere(got!=want)
    print('tuning',tuning,'nbad',len(bad), bad[:10].tolist())
    co=to_oracle(tr).reverse(o.pixel_grid(ny,nx))
    for y,x in bad[:10]: print((y,x), co[y,x], got[y,x], want[y,x])
