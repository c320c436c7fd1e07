#!/bin/bash
# diagnostics for the fused attention kernel: error maps on small cases, hang bisection on larger ones
for spec in "1 6 6 0" "1 6 18 0" "1 6 18 3" "1 12 12 0" "2 12 30 0" "3 60 66 0" "5 120 126 0" "4 240 240 0"; do
timeout -s KILL 40 python - $spec <<'PY'
import sys
sys.path.insert(0, 'waifu2x-tensorrt_b200'); sys.path.insert(0, 'tests')
import numpy as np, w2x
from test_gpu_swin_attn import make_case, reference
n, h, w, shift = map(int, sys.argv[1:5])
case = make_case(n, h, w, 5)
out, _ = w2x.run_swin_attn(*case, shift=shift)
ref = reference(*case, 6, shift)
err = np.abs(out.astype(np.float32) - ref)
print('case', n, h, w, shift, 'max', err.max(), 'mean', err.mean(), flush=True)
if err.max() > 0.03 and n * h * w <= 200:
    e = err.reshape(-1, 96)
    print(' per token max:', np.round(e.max(1), 2).tolist())
    print(' per 16-channel group max:', np.round(e.reshape(-1, 6, 16).max((0, 2)), 2).tolist())
elif err.max() > 0.03:
    e = err.reshape(n, h // 6, 6, w // 6, 6, 96).transpose(0, 1, 3, 2, 4, 5).reshape(-1, 36, 96)
    bad = (e.max((1, 2)) > 0.03)
    print(' bad windows: %d of %d; first bad ids %s' % (bad.sum(), bad.size, np.nonzero(bad)[0][:20].tolist()))
    print(' per position max (bad windows):', np.round(e[bad].max((0, 2)), 2).tolist())
PY
echo "rc=$?"
done
