import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from tests.test_gpu_training import _setup, _rel
from oracle import x3d_oracle as O, x3d_train_oracle as TO
from x3d_tf_b200 import training as TR
cfg, W, x, labels, mask, tr = _setup("X3D_XS", s=64, dropout=0.0)
ref = TO.train_step(W, O.OracleSpec.from_cfg(cfg), x, labels, lr=0.05, weight_decay=float(cfg.NETWORK.WEIGHT_DECAY), dropout_mask=mask, taps=(taps:={}))
rec = {}
orig = TR.X3DTrainer._bn_fwd
def patched(self, x2d, prefix, relu, tape):
    y = orig(self, x2d, prefix, relu, tape)
    inner = tape[-1]
    def bwd(dy):
        rec[prefix] = (x2d, y, relu, dy.clone())
        return inner(dy)
    tape[-1] = bwd
    return y
TR.X3DTrainer._bn_fwd = patched
loss = tr.step(torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), 0.05)
G = tr.grads()
for pfx in ["conv5/layer_with_weights-1", "stages/3/stage/layer_with_weights-6/bottleneck/bn_c"]:
    x2d, y, relu, dy = rec[pfx]
    g = dy.double().view(x2d.shape) * ((y > 0).double() if relu else 1.0)
    dbeta = g.sum(0).cpu().numpy()
    c = ref["grads"][pfx + "/beta"].shape[0]
    print(pfx, "recomputed-vs-kernel", _rel(dbeta[:c], G[pfx + "/beta"]), "recomputed-vs-oracle", _rel(dbeta[:c], ref["grads"][pfx + "/beta"]),
          "dy shape", tuple(dy.shape), "x shape", tuple(x2d.shape), "dy sum", float(dy.double().sum()))

x2d, y, relu, dy = rec["conv5/layer_with_weights-1"]
t = taps["conv5"]          # NCDHW
want_y = t.detach().permute(0,2,3,4,1).reshape(-1, t.shape[1]).numpy()
want_dy = t.grad.permute(0,2,3,4,1).reshape(-1, t.shape[1]).numpy()
print("a5 fwd err", _rel(y.cpu().numpy(), want_y), "dy err", _rel(dy.cpu().numpy(), want_dy), "mask mismatches", int(((y.cpu().numpy()>0) != (want_y>0)).sum()))
print("pool grad:", _rel((dy.view(2,-1,432).sum(1)).cpu().numpy(), taps["pool5"].grad.reshape(2,432).numpy()))
