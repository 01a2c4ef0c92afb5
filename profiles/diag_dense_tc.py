import sys, numpy as np, torch
sys.path.insert(0, ".")
from tests import helpers as H
from oracle import gemnet_oracle as go
from gemnet_pytorch_b200 import GemNet, build_batch, engine, dense_ops
dev = torch.device("cuda")
fx = H.load_npz("model_T.npz"); kw = H.model_kwargs(fx)
sd = go.synthetic_state_dict(kw, seed=1234)
model = GemNet(**kw); model.load_state_dict(sd); model = model.to(dev).eval()
inputs = {k: v.to(dev) for k, v in build_batch(*H.molecules(fx["mol_ids"].tolist()), 5.0, 10.0, True).items()}
def err(tag):
    E, F = model(inputs)
    dE = np.abs(E.double().cpu().numpy() - fx["E64"]).max() / np.abs(fx["E64"]).max()
    dF = np.abs(F.double().cpu().numpy() - fx["F64"]).max() / np.abs(fx["F64"]).max()
    print(f"{tag:40s} dE {dE:.2e} dF {dF:.2e}")
model.dense_algo = 0; err("cuBLAS")
model.dense_algo = 1; err("dense_tc all")
R = engine._Run
orig = (R.linear, R.dense, R.residual)
# bisect: disable TC in one of the three entry points at a time
def no_tc(fn):
    def w(self, *a, **k):
        tc = self.tc; self.tc = False
        try: return fn(self, *a, **k)
        finally: self.tc = tc
    return w
R.linear = no_tc(orig[0]); err("TC: dense+residual (linear via cuBLAS)"); R.linear = orig[0]
R.dense = no_tc(orig[1]); err("TC: linear+residual (dense via cuBLAS)"); R.dense = orig[1]
R.residual = no_tc(orig[2]); err("TC: linear+dense (residual via cuBLAS)"); R.residual = orig[2]
