import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from tests.test_gpu_training import _setup, _rel
from oracle import x3d_oracle as O, x3d_train_oracle as TO
S=int(sys.argv[2]) if len(sys.argv)>2 else 64
cfg, W, x, labels, mask, tr = _setup("X3D_XS", s=S, dropout=float(sys.argv[1]) if len(sys.argv)>1 else 0.5)
ref = TO.train_step(W, O.OracleSpec.from_cfg(cfg), x, labels, lr=0.05, weight_decay=float(cfg.NETWORK.WEIGHT_DECAY), dropout_mask=mask)
loss = tr.step(torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), 0.05)
print('logits rel err', _rel(tr.last_logits.cpu().numpy(), ref['logits']), 'loss', float(loss.mean()), ref['loss'])
G = tr.grads()
for k in list(tr.layout.slots)[::-1]:
    if k not in ref["grads"]: continue
    got = G[k].astype(np.float64)
    if TO.is_regularised(k): got = got + 2*float(cfg.NETWORK.WEIGHT_DECAY)*W[k]
    print(f"{_rel(got, ref['grads'][k]):9.2e}  {np.abs(ref['grads'][k]).max():9.2e}  {k}")
