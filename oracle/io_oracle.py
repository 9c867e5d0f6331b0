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
