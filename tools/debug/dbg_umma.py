import sys, numpy as np
sys.path.insert(0, '.')
from aind_smartspim_destripe_b200 import engine as E, synthetic as S
shape=(2048,2048); sigma=128
img=S.synthetic_plane(*shape, seed=2)
p=E.make_params(dict(level=None,sigma=sigma,max_threshold=12))
def run(umma):
    eng=E.DestripeEngine(*shape,max_planes=2)
    eng.set_umma(umma)
    eng.set_debug_stop(E.STAGE_FILTER)
    eng.filter_chunk(img[None],p,out_dtype=np.float32)
    r=[eng.debug_fetch(E.FETCH_CH,l,1)[0] for l in range(1,5)]
    eng.close(); return r
ref=run(False)
for rep in range(6):
    got=run(True)
    for l,(a,b) in enumerate(zip(got,ref),1):
        d=np.abs(a-b); bad=d>1e-5*np.abs(b).max()+1e-7
        if bad.any():
            rows=np.unique(np.nonzero(bad)[0]); cols=np.nonzero(bad.any(0))[0]
            print(f"rep {rep} level {l}: max {d.max():.3e} bad rows {len(rows)} [{rows[:6]}..{rows[-3:]}] cols {cols.min()}..{cols.max()} ({len(cols)})")
        else: print(f"rep {rep} level {l}: ok max {d.max():.2e}")
