"""TF tensor-bundle reader: real table file shipped by the reference + known-answer CRCs
 + writer round trip."""
import os
import struct

import numpy as np
import pytest

from x3d_tf_b200 import tf_bundle as tb

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_crc32c_known_answers():
    assert tb.crc32c(b"123456789") == 0xE3069283            # standard CRC-32C check value
    assert tb.crc32c(b"") == 0
    # masked CRCs recovered from the shipped indices
    assert tb.mask_crc(tb.crc32c(struct.pack("<f", 0.9))) == 0xDFC7EBFD      # optimizer/momentum
    assert tb.mask_crc(tb.crc32c(struct.pack("<f", 0.0))) == 0x3A117BA6      # optimizer/decay
    for it, want in ((1088505, 0x0F59E96B), (938240, 0x7AED0249), (1876480, 0x948DF8A0)):
        assert tb.mask_crc(tb.crc32c(struct.pack("<q", it))) == want         # optimizer/iter
    for v in (0, 1, 0xDEADBEEF, 0xFFFFFFFF):
        assert tb.unmask_crc(tb.mask_crc(v)) == v
    # incremental == one-shot
    a, b = os.urandom(100), os.urandom(37)
    assert tb.crc32c(b, tb.crc32c(a)) == tb.crc32c(a + b)


def test_reads_shipped_index(tmp_path, checkpoint_index):
    prefix = str(tmp_path / "model")
    with open(os.path.join(GOLDEN, "X3D-M.model.index"), "rb") as f, \
            open(prefix + ".index", "wb") as g:
        g.write(f.read())
    rd = tb.BundleReader(prefix)                 # verifies every block checksum
    assert rd.num_shards == 1 and len(rd.keys()) == 789   # 476 model + 308 slots + 4 optimizer + object graph
    want = checkpoint_index["keys"]
    crcs = checkpoint_index["crc"]["X3D_M"]
    for (k, dt, shape, off, size), crc in zip(want, crcs):
        e = rd.entries[k]
        assert (e.dtype, list(e.shape), e.offset, e.size, e.crc32c) == (dt, shape, off, size, crc)
    e = rd.entries["conv1/conv_s/kernel" + tb.VAR_SUFFIX]
    assert (e.offset, e.size, e.crc32c) == (6817364, 2592, 0xBD120AC9)
    e = rd.entries["fc2/kernel" + tb.VAR_SUFFIX]
    assert (e.offset, e.size, e.crc32c) == (3538944, 3276800, 0x0526337F)
    assert max(x.offset + x.size for x in rd.entries.values()) == 30355274
    # the data shard is absent upstream: reading a tensor must fail loudly, not return garbage
    with pytest.raises(FileNotFoundError):
        rd.tensor("fc2/bias" + tb.VAR_SUFFIX)


def test_corrupt_index_detected(tmp_path):
    raw = bytearray(open(os.path.join(GOLDEN, "X3D-M.model.index"), "rb").read())
    raw[1000] ^= 0x40
    p = tmp_path / "m.index"
    p.write_bytes(bytes(raw))
    with pytest.raises(tb.BundleError):
        tb.BundleReader(str(tmp_path / "m"))
    p.write_bytes(bytes(raw[:-8]) + b"\0" * 8)
    with pytest.raises(tb.BundleError):
        tb.BundleReader(str(tmp_path / "m"))


def test_write_read_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {f"stages/{i}/stage/layer_with_weights-{j}/bottleneck/a/kernel":
               rng.normal(size=(1, 1, 1, 8 + i, 16 + j)).astype(np.float32)
               for i in range(3) for j in range(12)}
    tensors["fc2/bias"] = rng.normal(size=(400,)).astype(np.float32)
    tensors["optimizer/iter"] = np.array(1234567, np.int64)
    prefix = str(tmp_path / "ckpt" / "model")
    tb.write_bundle(prefix, tensors, block_size=512)      # small blocks: many data blocks
    assert tb.latest_checkpoint(str(tmp_path / "ckpt")) == prefix
    rd = tb.BundleReader(prefix)
    assert rd.keys() == sorted([k + tb.VAR_SUFFIX for k in tensors] + [tb.OBJECT_GRAPH_KEY])
    for k, v in tensors.items():
        got = rd.tensor(k + tb.VAR_SUFFIX)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
    mv = tb.load_model_variables(prefix)
    assert "optimizer/iter" not in mv and len(mv) == len(tensors) - 1
    # flip one data byte -> per-tensor CRC catches it
    path = rd.shard_path(0)
    raw = bytearray(open(path, "rb").read())
    raw[rd.entries["fc2/bias" + tb.VAR_SUFFIX].offset + 5] ^= 1
    open(path, "wb").write(bytes(raw))
    with pytest.raises(tb.BundleError):
        tb.BundleReader(prefix).tensor("fc2/bias" + tb.VAR_SUFFIX)
    assert tb.BundleReader(prefix).tensor("fc2/bias" + tb.VAR_SUFFIX, verify=False).shape == (400,)


def test_scalar_and_empty_dim_entries():
    e = tb.BundleEntry(dtype=tb.DT_INT64, shape=(), offset=6817356, size=8, crc32c=0x948DF8A0)
    assert tb.BundleEntry.parse(e.serialize()) == e
    e = tb.BundleEntry(dtype=tb.DT_FLOAT, shape=(3, 0, 2), offset=0, size=0, crc32c=1)
    assert tb.BundleEntry.parse(e.serialize()) == e


def test_optimizer_state_round_trip_and_tf_known_answers(tmp_path):
    """The optimizer part of a Keras checkpoint (utils.py:128-132): our writer
    must produce the entries the shipped X3D-M index holds for the same values (masked CRC-32C
    known answers), and the reader must give the slots back under their variable paths."""
    from x3d_tf_b200 import tf_bundle as B
    rng = np.random.default_rng(0)
    weights = {"conv1/conv_s/kernel": rng.normal(size=(1, 3, 3, 3, 24)).astype(np.float32),
               "conv1/bn/moving_mean": rng.normal(size=24).astype(np.float32)}
    slots = {"conv1/conv_s/kernel": rng.normal(size=(1, 3, 3, 3, 24)).astype(np.float32)}
    tensors = dict(weights)
    tensors.update(B.optimizer_tensors(1876480, 0.05, 0.9, slots))
    prefix = str(tmp_path / "ckpt-7")
    B.write_bundle(prefix, tensors)
    rd = B.BundleReader(prefix)
    entries = {k: B.BundleEntry.parse(v) for k, v in B.read_index(prefix + ".index").items() if k}
    sfx = B.VAR_SUFFIX
    assert entries["optimizer/momentum" + sfx].crc32c == 0xdfc7ebfd          # float32 0.9
    assert entries["optimizer/decay" + sfx].crc32c == 0x3a117ba6             # float32 0.0
    assert entries["optimizer/iter" + sfx].crc32c == 0x948df8a0              # int64 1 876 480 (X3D-M)
    assert "conv1/conv_s/kernel/.OPTIMIZER_SLOT/optimizer/momentum" + sfx in rd
    st = B.load_optimizer_state(prefix)
    assert st["iter"] == 1876480 and abs(st["momentum"] - 0.9) < 1e-7 and abs(st["learning_rate"] - 0.05) < 1e-8
    assert list(st["slots"]) == ["conv1/conv_s/kernel"]
    assert np.array_equal(st["slots"]["conv1/conv_s/kernel"], slots["conv1/conv_s/kernel"])
    # model variables are still read without the optimizer entries (expect_partial semantics)
    got = B.load_model_variables(prefix)
    assert sorted(got) == sorted(weights) and np.array_equal(got["conv1/bn/moving_mean"], weights["conv1/bn/moving_mean"])
    assert B.latest_checkpoint(str(tmp_path)).endswith("ckpt-7")


def test_train_driver_host_logic(tmp_path):
    """Epoch from the checkpoint name (train.py:133), the LR schedule (train.py:114-125) and the
    shuffled file batches of the training driver."""
    import math
    from x3d_tf_b200 import train as T
    from x3d_tf_b200.config import get_config
    assert T.epoch_of_checkpoint("/a/b/ckpt-17") == 17
    cfg = get_config("X3D_M")
    t = cfg.TRAIN
    for e in (0, 1, t.WARMUP_EPOCHS, t.WARMUP_EPOCHS + 1, 100, t.EPOCHS - 1):
        want = (t.BASE_LR * 0.5 * (math.cos(math.pi * e / t.EPOCHS) + 1) if e > t.WARMUP_EPOCHS
                else t.WARMUP_LR + e * (t.BASE_LR - t.WARMUP_LR) / t.WARMUP_EPOCHS)
        assert abs(T.lr_for_epoch(cfg, e) - want) < 1e-12
    items = []
    for i in range(5):
        np.save(tmp_path / f"c{i}.npy", np.full((2, 1, 2, 2, 3) if i % 2 else (1, 2, 2, 3), i, np.uint8))
        items.append((str(tmp_path / f"c{i}.npy"), i))
    batches = list(T.file_clip_batches(items, 2, 2, seed=0))
    assert len(batches) == 2 and all(b[0].shape == (2, 1, 2, 2, 3) for b in batches)
    for clips, labels in batches:                          # the label travels with its clip
        assert [int(c.flat[0]) for c in clips] == labels.tolist()
    syn = list(T.synthetic_clip_batches(2, 2, 1, 4, 400, 3))
    assert len(syn) == 2 and syn[0][0].dtype == np.uint8 and syn[0][1].dtype == np.int32
    # Every rank runs the SAME number of steps whatever the list length (a 5-item list split over 2
    # ranks used to give 3 / 2 items, i.e. different step counts and mismatched collectives), the
    # ranks' slices of one global batch are disjoint, and a short list repeats (Keras steps_per_epoch).
    per_rank = [list(T.file_clip_batches(items, 4, 3, seed=5, rank=r, world=2)) for r in range(2)]
    assert [len(b) for b in per_rank] == [3, 3]
    idx = list(T.global_batch_indices(5, 4, 3, seed=5))
    assert all(len(set(i.tolist())) == 4 for i in idx[:1]) and sum(len(i) for i in idx) == 12
    for step in range(3):
        l0, l1 = per_rank[0][step][1].tolist(), per_rank[1][step][1].tolist()
        assert l0 + l1 == idx[step].tolist()


def test_from_scratch_init_is_keras_default():
    """train.py:128 builds `X3D(cfg)` with Keras default initialisers: BN gamma=1, beta=0, mean=0,
    variance=1, zero biases, Glorot-uniform kernels with `_compute_fans` of the kernel tensor
    (a channelwise (3,3,3,1,C) kernel: fan_in 27, fan_out 27*C)."""
    from x3d_tf_b200.config import get_config
    from x3d_tf_b200.model import keras_default_weights
    W = keras_default_weights(get_config("X3D_M"), seed=1111)
    assert len(W) == 476
    for k, v in W.items():
        leaf = k.rsplit("/", 1)[1]
        if leaf in ("gamma", "moving_variance"):
            assert np.all(v == 1.0), k
        elif leaf in ("beta", "moving_mean", "bias"):
            assert np.all(v == 0.0), k
        else:
            shp = v.shape
            rf = int(np.prod(shp[:-2])) if len(shp) == 5 else 1
            lim = np.sqrt(6.0 / (rf * shp[-2] + rf * shp[-1]))
            assert np.abs(v).max() <= lim + 1e-7, k
            if v.size > 500:
                assert np.abs(v).max() > 0.9 * lim, k
    b = W["stages/0/stage/layer_with_weights-0/bottleneck/b/kernel"]
    assert b.shape == (3, 3, 3, 1, 54) and np.abs(b).max() <= np.sqrt(6.0 / (27 + 27 * 54)) + 1e-7
    W2 = keras_default_weights(get_config("X3D_M"), seed=1111)
    assert all(np.array_equal(W[k], W2[k]) for k in W)


def test_written_checkpoint_carries_a_keras_object_graph(tmp_path, checkpoint_index, B=tb):
    """`write_bundle` emits `_CHECKPOINTABLE_OBJECT_GRAPH` (utils.py:128-132 checkpoints, read back by
    train.py:137 / eval.py:81 through Keras' object-based restore): a DT_STRING scalar in TF's string
    framing whose TrackableObjectGraph reaches EVERY variable of the shipped X3D-M index by walking
    `local_name` edges along its checkpoint key, with the momentum slots hung off the optimizer node."""
    from x3d_tf_b200.arch import build_arch
    from x3d_tf_b200.config import get_config
    from x3d_tf_b200.synth import synthetic_weights
    W = synthetic_weights(build_arch(get_config("X3D_M")), seed=1)
    slots = {k: np.zeros_like(v) for k, v in W.items()
             if not k.endswith(("moving_mean", "moving_variance"))}
    tensors = dict(W)
    tensors.update(B.optimizer_tensors(1876480, 0.05, 0.9, slots))
    prefix = str(tmp_path / "ckpt-3")
    B.write_bundle(prefix, tensors)
    rd = B.BundleReader(prefix)
    suffix = "/.ATTRIBUTES/VARIABLE_VALUE"
    shipped = {e[0] for e in checkpoint_index["keys"]}
    assert set(rd.keys()) == shipped                          # incl. the object graph and the "" header
    e = rd.entries[B.OBJECT_GRAPH_KEY]
    assert e.dtype == B.DT_STRING and e.shape == ()
    (proto,) = rd.string_tensor(B.OBJECT_GRAPH_KEY)
    assert e.size == len(proto) + 4 + len(B._put_varint(len(proto)))          # TF's string-tensor framing
    raw = bytes(rd._shard(0)[e.offset:e.offset + e.size])
    assert B.unmask_crc(e.crc32c) == B.crc32c(proto, B.crc32c(raw[len(raw) - len(proto) - 4:len(raw) - len(proto)],
                                                               B.crc32c(np.uint32(len(proto)).tobytes())))
    nodes = B.parse_object_graph(proto)
    var_keys = sorted(k for k in shipped if k.endswith(suffix))
    assert len(var_keys) == 788
    reached = set()
    opt = nodes[0]["children"]["optimizer"]
    slot_of = {(v, name): sid for v, name, sid in nodes[opt]["slot_variables"]}
    assert len(slot_of) == 308
    for k in var_keys:
        path = k[:-len(suffix)]
        if "/.OPTIMIZER_SLOT/" in path:
            var_path, rest = path.split("/.OPTIMIZER_SLOT/")
            assert rest == "optimizer/momentum"
            cur = 0
            for part in var_path.split("/"):
                cur = nodes[cur]["children"][part]
            nid = slot_of[(cur, "momentum")]
        else:
            nid = 0
            for part in path.split("/"):
                nid = nodes[nid]["children"][part]            # KeyError = Keras could not match the dependency
        (name, full, key), = nodes[nid]["attributes"]
        assert name == "VARIABLE_VALUE" and key == k
        reached.add(nid)
    assert len(reached) == 788
    # interior nodes carry no tensors; the root is node 0 and has the model's attribute names
    assert {"conv1", "stages", "conv5", "fc1", "fc2", "optimizer"} <= set(nodes[0]["children"])
    assert set(nodes[nodes[0]["children"]["stages"]]["children"]) == {"0", "1", "2", "3"}
    # model variables still load by name, the graph entry is ignored there
    got = B.load_model_variables(prefix)
    assert sorted(got) == sorted(W)
