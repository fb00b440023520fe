import numpy as np, sys
sys.path.insert(0,'.')
from ramscb_b200 import grids, synthetic, host
from oracle import oracle
g=grids.build_grids()
inp=synthetic.make_inputs(g, f2_kind="noisy", inductive=True)
o=oracle.RamOracle(g, inp, DTs=5.0)
gpu=host.RamGpu(g); gpu.set_inputs(inp)
o.set_scalar("T",10.0)
S=1
o.op("coulpara",S); gpu.COULPARA(S,5.0)
o.op("coulen",S); gpu.COULEN(S)
got=gpu.f2_d2h()[S-1]; ref=o.F2[S-1]
bad=np.argwhere(~(got==ref))
print("n bad", len(bad), "nan", np.isnan(got).sum(), "inf", np.isinf(got).sum())
print(bad[:10])
for b in bad[:5]:
    print(tuple(b), got[tuple(b)], ref[tuple(b)], inp.F2[S-1][tuple(b)])
import collections
print(collections.Counter(bad[:,3]).most_common(5), collections.Counter(bad[:,2]).most_common(5))
