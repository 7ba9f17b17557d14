"""Mean time of the bare row-tile RHS kernel (QSX_HEOM_TIME diagnostic of qsx_heom_apply) for
several environments.  usage: python tools/apply_time.py [depth] [KEY=VAL,KEY=VAL ...] ..."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch, qspectra_b200 as qb
from qspectra_b200 import systems
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
specs = sys.argv[2:] or ['QSX_HEOM_GRID=296', 'QSX_HEOM_GRID=148']
model = qb.HEOMModel(systems.fmo(), hilbert_subspace='e', unit_convert=qb.CM_FS, level_cutoff=depth, K=1)
eom = model.equation_of_motion('ee')
rng = np.random.RandomState(0)
y = torch.from_numpy(rng.randn(1, eom.dim) + 1j * rng.randn(1, eom.dim)).cuda()
dy = torch.empty_like(y)
ref = None
for spec in specs:
    for k in [k for k in os.environ if k.startswith('QSX_HEOM_')]:
        os.environ.pop(k)
    for kv in spec.split(','):
        k, v = kv.split('=')
        os.environ[k] = v
    os.environ['QSX_HEOM_TIME'] = '20'
    print(spec, flush=True)
    eom._apply_dev(y, dy, 1, None)
    eom._apply_dev(y, dy, 1, None)
    out = dy.cpu().numpy()
    if ref is None:
        ref = out
    print('   rel diff vs first %.2e' % (np.linalg.norm(out - ref) / np.linalg.norm(ref)), flush=True)
