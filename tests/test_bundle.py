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
    assert rd.keys() == sorted(k + tb.VAR_SUFFIX for k in tensors)
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
