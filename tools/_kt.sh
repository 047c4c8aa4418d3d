python -m pytest tests/test_gpu_batched_affine.py -x -q 2>&1 | tail -3
python tools/ba_tune.py 20 2>&1 | tail -14
cat > /tmp/one.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import torch
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy
n = 1 << 20
ctx = zk.Context(0); ctx.set_opt(1, 0); ctx.set_opt(2, 3)
bases = zk.scalar_mul_many(ctx, 1, zk.G1_GENERATOR, sy.random_fr_limbs(n, 1))
b = zk.Bases(ctx, 1, bases, precompute=True)
d = torch.from_numpy(sy.random_fr_limbs(n, 2).view(np.int64)).cuda(); torch.cuda.synchronize()
for _ in range(3): zk.multiexp_device(b, d.data_ptr(), n)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_l_v3.csv python /tmp/one.py > /dev/null 2>&1
