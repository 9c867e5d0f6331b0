import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from tests.test_gpu_training import _setup, _rel, _l2
from oracle import x3d_oracle as O, x3d_train_oracle as TO
cfg, W, x, labels, mask, tr = _setup("X3D_XS", dropout=0.0)
spec = O.OracleSpec.from_cfg(cfg); wd = float(cfg.NETWORK.WEIGHT_DECAY)
r1 = TO.train_step(W, spec, x, labels, lr=1e-4, weight_decay=wd)
W1 = {k: v.astype(np.float32) for k, v in r1["weights"].items()}
r2 = TO.train_step(W1, spec, x, labels, lr=5e-5, weight_decay=wd, velocity=r1["velocity"])
xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda()
tr.step(xd, ld, 1e-4)
Wa = tr.weights()
tr.step(xd, ld, 5e-5)
Wb = tr.weights()
for k in ["fc2/kernel", "fc2/bias", "fc1/kernel", "conv1/conv_s/kernel", "conv5/layer_with_weights-0/kernel"]:
    d1t, d1o = Wa[k].astype(np.float64) - W[k], r1["weights"][k] - W[k]
    d2t, d2o = Wb[k].astype(np.float64) - Wa[k], r2["weights"][k] - W1[k]
    print(k, "step1 l2", _l2(d1t, d1o), "norm ratio", np.linalg.norm(d1t)/np.linalg.norm(d1o), "| step2 l2", _l2(d2t, d2o), "ratio", np.linalg.norm(d2t)/np.linalg.norm(d2o))
G2t = tr.grads()
Wa32 = {k: v.astype(np.float32) for k, v in Wa.items()}
ra = TO.train_step(Wa32, spec, x, labels, lr=5e-5, weight_decay=wd)
for k in ["fc2/kernel", "fc1/kernel", "conv1/conv_s/kernel", "conv5/layer_with_weights-0/kernel"]:
    gt = G2t[k].astype(np.float64) + (2*wd*Wa[k] if TO.is_regularised(k) else 0)
    print(k, "trainer g2 vs oracle@Wa", _l2(gt, ra["grads"][k]), "oracle@W1 vs oracle@Wa", _l2(r2["grads"][k], ra["grads"][k]), "g1 vs g2 (oracle)", _l2(r1["grads"][k], r2["grads"][k]))
