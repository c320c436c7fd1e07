import sys; sys.path.insert(0,"."); sys.path.insert(0,"waifu2x-tensorrt_b200")
import __graft_entry__, w2x, tempfile
import numpy as np
from oracle import tiling
d=tempfile.mkdtemp(); m,p=__graft_entry__.make_synthetic_model(d, scale=2)
e=w2x.Img2Img(); e.setMessageCallback(lambda s,m: print("MSG",s,m))
print(e.build(p, w2x.BuildConfig.fixed(4,64)), e.load(p, w2x.RenderConfig(batchSize=4,height=64,width=64,scaling=2)))
src=tiling.synthetic_frame(150,130,3); r=e.render(src); print(r is not None)
g=w2x.render_banded([e], src); print(g is not None, e.last_error)
if g is not None:
    d=np.abs(g.astype(int)-r.astype(int)); print("maxdiff", d.max(), "ndiff", (d>0).sum(), "of", d.size)
