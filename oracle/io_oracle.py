"""CPU restatement (numpy) of the two steps either side of the forward path.  TEST INFRASTRUCTURE
ONLY: imported by tests/, never by x3d_tf_b200/ (see oracle/x3d_oracle.py header).

  * normalize        -- reference utils.py:42-72 (`utils.normalize`, called by dataloader.py on
                        decoded frames): float32 arithmetic, x / norm_value, then (x - mean) / std.
  * eval_metrics     -- what eval.py:62-70 compiles into the model and eval.py:83 evaluates.  The
                        arithmetic lives in tensorflow==2.4.1 (requirements.txt:4, absent here):
      - loss: keras.backend.sparse_categorical_crossentropy(from_logits=False) on a tensor that is
        not the direct output of a Softmax op (eval-mode X3D.call returns a reduce_mean over
        views, model.py:123-127): output = clip(output, 1e-7, 1 - 1e-7); then
        sparse_softmax_cross_entropy_with_logits(labels, log(output))
        = log(sum_j output_j) - log(output_label); mean over the batch.
      - 'acc': SparseCategoricalAccuracy = (argmax(y_pred) == label), first index on ties.
      - 'top_5_acc': SparseTopKCategoricalAccuracy(k=5) = tf.math.in_top_k: the label is in the
        top k iff fewer than k classes have a strictly larger prediction.
    Parity unpinned against real TensorFlow (none of this is executable here); pinned against
    hand-worked cases in tests/test_io_oracle.py.
"""
from __future__ import annotations

import numpy as np


def normalize(clips_u8: np.ndarray, mean, std, norm_value: float = 255.0) -> np.ndarray:
    x = clips_u8.astype(np.float32)
    x = x / np.float32(norm_value)
    x = x - np.asarray(mean, np.float32)
    return (x / np.asarray(std, np.float32)).astype(np.float32)


def eval_metrics(probs: np.ndarray, labels: np.ndarray, k: int = 5) -> dict:
    probs = np.asarray(probs, np.float32)
    labels = np.asarray(labels).reshape(-1)
    V = probs.shape[0]
    eps = np.float32(1e-7)
    clipped = np.clip(probs, eps, np.float32(1.0) - eps)
    picked = clipped[np.arange(V), labels]
    loss = np.log(clipped.sum(1, dtype=np.float32)) - np.log(picked)
    top1 = probs.argmax(1) == labels
    pl = probs[np.arange(V), labels][:, None]
    topk = (probs > pl).sum(1) < k
    return {"loss": float(loss.astype(np.float64).mean()), "acc": float(top1.mean()),
            f"top_{k}_acc": float(topk.mean()), "videos": V,
            "sums": np.array([loss.astype(np.float64).sum(), top1.sum(), topk.sum(), V], np.float64)}


def eval_views(video: np.ndarray, T: int, views: int, crops: int, size: int) -> np.ndarray:
    """Evaluation clips of one decoded, resized video [F,H,W,C] as the reference builds them
    (parity unpinned against TensorFlow; restated line by line):
      * temporal (transforms.py:48-65): sample_rate = max(1, F // T); end = T * sample_rate * views;
        indices = tile(range(F), loops)[0:end][0:end:sample_rate]; reshape to [views, T, ...];
      * spatial (transforms.py:149-190, 216-222): crop i of `crops` uses spatial_idx = i % 3 if
        crops > 1 else 1; offsets ceil((dim - size) / 2), the longer side gets 0 / dim - size for
        idx 0 / 2 (`if height > width` ... `else` ...);
      * batching (dataloader.py:107-116): [crops, views, T, S, S, C] reshaped to [-1, T, S, S, C]."""
    F, H, W, _ = video.shape
    rate = max(1, F // T)
    end = T * rate * views
    loops = -(-end // F)
    idx = np.tile(np.arange(F), loops)[0:end][0:end:rate]
    clip = video[idx].reshape(views, T, H, W, video.shape[3])
    out = []
    for i in range(crops):
        sidx = i % 3 if crops > 1 else 1
        y0 = int(np.ceil((H - size) / 2))
        x0 = int(np.ceil((W - size) / 2))
        if H > W:
            if sidx == 0:
                y0 = 0
            elif sidx == 2:
                y0 = H - size
        else:
            if sidx == 0:
                x0 = 0
            elif sidx == 2:
                x0 = W - size
        out.append(clip[:, :, y0:y0 + size, x0:x0 + size, :])
    return np.stack(out).reshape(-1, T, size, size, video.shape[3])
