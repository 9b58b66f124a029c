import sys, torch
sys.path.insert(0,'/root/repo')
import warnings; warnings.filterwarnings('ignore')
from tests import parity
dev=torch.device('cuda:0')
for case in ['city_x4','kitti_x2','train_lo']:
    cfg,_,z = parity.load_case(case)
    o32,g32 = parity.oracle_decode(cfg,z,torch.float32)
    o64,g64 = parity.oracle_decode(cfg,z,torch.float64)
    out,g = parity.kernel_decode(dev,cfg,z)
    print(case)
    for n,o,a in zip(parity.OUT_NAMES,out,o32):
        if a is None: continue
        print('  %-10s %.2e' % (n, float((o.cpu()-a).abs().max())))
    for k in parity.LEAF_KEYS:
        gm = float(g64[k].abs().max())
        print('  grad %-6s k-ref32 %.2e  ref32-f64 %.2e (max %.2e)' % (k, float((g[k].cpu()-g32[k]).abs().max())/gm, float((g32[k]-g64[k]).abs().max())/gm, gm))
