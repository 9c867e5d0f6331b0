"""Reader (and minimal writer) for TensorFlow "tensor bundle" checkpoints.

The reference restores weights with `model.load_weights(prefix)` (`train.py:137-143`,
`eval.py:78-81`) from bundles such as `models/X3D-M/{checkpoint,model.index,
model.data-00000-of-00001}`.  TensorFlow is not available here, so the on-disk format is read
directly:

  * `<prefix>.index` is a LevelDB-style sorted string table: data blocks of prefix-compressed
    (key, value) entries, an index block of block handles, and a 48-byte footer ending in the
    magic 0xdb4775248b80fb57.  Every block is followed by a 1-byte compression tag and a masked
    CRC-32C.
  * key ""  -> BundleHeaderProto {1:num_shards, 2:endianness, 3:version}
  * key k   -> BundleEntryProto  {1:dtype, 2:shape{2:dim{1:size}}, 3:shard_id, 4:offset,
                                  5:size, 6:crc32c (fixed32, masked)}
  * `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian row-major tensor bytes.

Model variables are addressed directly by their object-graph attribute path
(`conv1/conv_s/kernel/.ATTRIBUTES/VARIABLE_VALUE`, ...), which is what Keras' `load_weights`
resolves to for this model; optimizer slots and the object-graph string are ignored, like
`expect_partial()` does (`eval.py:81`).
"""
from __future__ import annotations

import os
import re
import struct
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
VAR_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

DT_FLOAT, DT_DOUBLE, DT_INT32, DT_STRING, DT_INT64, DT_BOOL, DT_BFLOAT16, DT_HALF = \
    1, 2, 3, 7, 9, 10, 14, 19
_NP_OF_DT = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"),
             DT_INT64: np.dtype("<i8"), DT_BOOL: np.dtype("bool"), DT_HALF: np.dtype("<f2")}
_DT_OF_NP = {v: k for k, v in _NP_OF_DT.items()}


class BundleError(IOError):
    pass


# ------------------------------------------------------------------------------ CRC-32C
_CRC_TABLE: Optional[List[int]] = None
_native_crc = None


def _table() -> List[int]:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    return _CRC_TABLE


def set_native_crc32c(fn) -> None:
    """Install a fast CRC-32C (the C-ABI library exports `x3d_crc32c`); optional."""
    global _native_crc
    _native_crc = fn


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C (Castagnoli), unmasked."""
    if _native_crc is not None and len(data) >= 64:
        return _native_crc(data, crc)
    t = _table()
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(m: int) -> int:
    rot = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------------------ varints / protobuf
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        if pos >= len(buf):
            raise BundleError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise BundleError("varint too long")


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_fields(buf: bytes) -> Iterable[Tuple[int, int, object]]:
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        fno, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise BundleError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


@dataclass
class BundleEntry:
    dtype: int = 0
    shape: Tuple[int, ...] = ()
    shard_id: int = 0
    offset: int = 0
    size: int = 0
    crc32c: int = 0          # masked, as stored
    has_slices: bool = False

    @classmethod
    def parse(cls, buf: bytes) -> "BundleEntry":
        e = cls()
        for fno, _, v in _pb_fields(buf):
            if fno == 1:
                e.dtype = v
            elif fno == 2:
                dims = []
                for f2, _, v2 in _pb_fields(v):
                    if f2 == 2:
                        size = 0
                        for f3, _, v3 in _pb_fields(v2):
                            if f3 == 1:
                                size = _signed64(v3)
                        dims.append(size)
                e.shape = tuple(dims)
            elif fno == 3:
                e.shard_id = v
            elif fno == 4:
                e.offset = v
            elif fno == 5:
                e.size = v
            elif fno == 6:
                e.crc32c = v
            elif fno == 7:
                e.has_slices = True
        return e

    def serialize(self) -> bytes:
        out = bytearray()
        if self.dtype:
            out += b"\x08" + _put_varint(self.dtype)
        shp = bytearray()
        for d in self.shape:
            dim = b"\x08" + _put_varint(d) if d else b""
            shp += b"\x12" + _put_varint(len(dim)) + dim
        out += b"\x12" + _put_varint(len(shp)) + bytes(shp)
        if self.shard_id:
            out += b"\x18" + _put_varint(self.shard_id)
        if self.offset:
            out += b"\x20" + _put_varint(self.offset)
        if self.size:
            out += b"\x28" + _put_varint(self.size)
        out += b"\x35" + struct.pack("<I", self.crc32c)
        return bytes(out)


# ------------------------------------------------------------------------------ table reader
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    if offset + size + 5 > len(data):
        raise BundleError("block handle out of range")
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if unmask_crc(stored) != crc32c(data[offset:offset + size + 1]):
            raise BundleError(f"index block at {offset}: checksum mismatch")
    if ctype != 0:
        raise BundleError(f"index block compression type {ctype} not supported (expected 0)")
    return contents


def _block_entries(block: bytes) -> Iterable[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise BundleError("block too small")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    if limit < 0:
        raise BundleError("bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key):
            raise BundleError("corrupt prefix compression")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(path: str, verify: bool = True) -> "OrderedDict[str, bytes]":
    """All (key, raw value) pairs of an `.index` table, in file (sorted) order."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48:
        raise BundleError(f"{path}: too small for a table footer")
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise BundleError(f"{path}: bad table magic")
    pos = 0
    _, pos = _get_varint(footer, pos)      # metaindex offset
    _, pos = _get_varint(footer, pos)      # metaindex size
    ioff, pos = _get_varint(footer, pos)
    isz, pos = _get_varint(footer, pos)
    out: "OrderedDict[str, bytes]" = OrderedDict()
    for _, handle in _block_entries(_read_block(data, ioff, isz, verify)):
        boff, p = _get_varint(handle, 0)
        bsz, p = _get_varint(handle, p)
        for k, v in _block_entries(_read_block(data, boff, bsz, verify)):
            out[k.decode("utf-8")] = v
    return out


class BundleReader:
    """Random access to the tensors of one checkpoint prefix."""

    def __init__(self, prefix: str, verify_index: bool = True):
        self.prefix = prefix
        raw = read_index(prefix + ".index", verify_index)
        if "" not in raw:
            raise BundleError("bundle header (empty key) missing")
        self.num_shards, self.endianness = 1, 0
        for fno, _, v in _pb_fields(raw[""]):
            if fno == 1:
                self.num_shards = v
            elif fno == 2:
                self.endianness = v
        if self.endianness != 0:
            raise BundleError("big-endian bundles are not supported")
        self.entries: "OrderedDict[str, BundleEntry]" = OrderedDict(
            (k, BundleEntry.parse(v)) for k, v in raw.items() if k != "")
        self._shards: Dict[int, np.memmap] = {}

    def keys(self) -> List[str]:
        return list(self.entries)

    def __contains__(self, key: str) -> bool:
        return key in self.entries

    def shard_path(self, shard_id: int) -> str:
        return f"{self.prefix}.data-{shard_id:05d}-of-{self.num_shards:05d}"

    def _shard(self, shard_id: int):
        if shard_id not in self._shards:
            path = self.shard_path(shard_id)
            if not os.path.exists(path):
                raise FileNotFoundError(
                    f"checkpoint data shard missing: {path} (the index lists "
                    f"{max(e.offset + e.size for e in self.entries.values())} bytes)")
            self._shards[shard_id] = np.memmap(path, dtype=np.uint8, mode="r")
        return self._shards[shard_id]

    def string_tensor(self, key: str) -> List[bytes]:
        """A DT_STRING entry (e.g. `_CHECKPOINTABLE_OBJECT_GRAPH`) as a list of byte strings."""
        e = self.entries[key]
        if e.dtype != DT_STRING:
            raise BundleError(f"{key}: not a string tensor")
        shard = self._shard(e.shard_id)
        raw = bytes(shard[e.offset:e.offset + e.size])
        n = int(np.prod(e.shape, dtype=np.int64)) if e.shape else 1
        return _parse_string_tensor(raw, n)

    def tensor(self, key: str, verify: bool = True) -> np.ndarray:
        e = self.entries[key]
        if e.has_slices:
            raise BundleError(f"{key}: sliced (partitioned) variables are not supported")
        if e.dtype not in _NP_OF_DT:
            raise BundleError(f"{key}: dtype enum {e.dtype} not supported")
        dt = _NP_OF_DT[e.dtype]
        n = int(np.prod(e.shape, dtype=np.int64)) if e.shape else 1
        if n * dt.itemsize != e.size:
            raise BundleError(f"{key}: size {e.size} does not match shape {e.shape}")
        shard = self._shard(e.shard_id)
        if e.offset + e.size > shard.shape[0]:
            raise BundleError(f"{key}: extends past the end of the data shard")
        raw = bytes(shard[e.offset:e.offset + e.size])
        if verify and unmask_crc(e.crc32c) != crc32c(raw):
            raise BundleError(f"{key}: tensor checksum mismatch")
        return np.frombuffer(raw, dtype=dt).reshape(e.shape).copy()


def latest_checkpoint(model_dir: str) -> Optional[str]:
    """`tf.train.latest_checkpoint`: read the `checkpoint` state file (text proto)."""
    state = os.path.join(model_dir, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state) as f:
        m = re.search(r'^model_checkpoint_path:\s*"(.*)"', f.read(), re.M)
    if not m:
        return None
    p = m.group(1)
    return p if os.path.isabs(p) else os.path.join(model_dir, p)


def load_model_variables(prefix: str, verify: bool = True,
                         names: Optional[Iterable[str]] = None) -> Dict[str, np.ndarray]:
    """{attribute path -> float32 array} for every model variable of the bundle (optimizer
    state, slot variables and the object graph are skipped -- `expect_partial`)."""
    rd = BundleReader(prefix)
    want = set(names) if names is not None else None
    out = {}
    for k in rd.keys():
        if not k.endswith(VAR_SUFFIX) or "/.OPTIMIZER_SLOT/" in k or k.startswith("optimizer/"):
            continue
        name = k[:-len(VAR_SUFFIX)]
        if want is not None and name not in want:
            continue
        out[name] = rd.tensor(k, verify)
    return out


SLOT_INFIX = "/.OPTIMIZER_SLOT/optimizer/momentum"
SLOT_M, SLOT_V = "/.OPTIMIZER_SLOT/optimizer/m", "/.OPTIMIZER_SLOT/optimizer/v"       # Adam


def load_optimizer_state(prefix: str, verify: bool = True) -> Dict[str, object]:
    """Optimizer part of a Keras checkpoint written by `ModelCheckpoint` (`utils.py:128-132`):
    `optimizer/{iter,learning_rate,momentum,decay}` and the SGD momentum slots
    `<variable>/.OPTIMIZER_SLOT/optimizer/momentum`.  Returns
    {"iter": int, "learning_rate": float|None, "momentum": float|None, "decay": float|None,
     "slots": {variable path -> float32 array}}; missing pieces are None / empty."""
    rd = BundleReader(prefix)
    out: Dict[str, object] = {"iter": None, "learning_rate": None, "momentum": None, "decay": None,
                              "beta_1": None, "beta_2": None, "slots": {}, "slots_m": {}, "slots_v": {}}
    for k in rd.keys():
        if not k.endswith(VAR_SUFFIX):
            continue
        name = k[:-len(VAR_SUFFIX)]
        if name.startswith("optimizer/"):
            v = rd.tensor(k, verify)
            field = name[len("optimizer/"):]
            if field == "iter":
                out["iter"] = int(np.asarray(v).reshape(-1)[0])
            elif field in out:
                out[field] = float(np.asarray(v).reshape(-1)[0])
        elif name.endswith(SLOT_INFIX):
            out["slots"][name[:-len(SLOT_INFIX)]] = rd.tensor(k, verify)
        elif name.endswith(SLOT_M):
            out["slots_m"][name[:-len(SLOT_M)]] = rd.tensor(k, verify)
        elif name.endswith(SLOT_V):
            out["slots_v"][name[:-len(SLOT_V)]] = rd.tensor(k, verify)
    return out


def adam_tensors(iteration: int, learning_rate: float, beta_1: float, beta_2: float,
                 m: Dict[str, np.ndarray], v: Dict[str, np.ndarray], decay: float = 0.0) -> Dict[str, np.ndarray]:
    """Keras Adam's checkpoint entries: `optimizer/{iter,learning_rate,beta_1,beta_2,decay}` and the
    `m` / `v` slots of every trainable variable."""
    out: Dict[str, np.ndarray] = {
        "optimizer/iter": np.asarray(iteration, np.int64),
        "optimizer/learning_rate": np.asarray(learning_rate, np.float32),
        "optimizer/beta_1": np.asarray(beta_1, np.float32),
        "optimizer/beta_2": np.asarray(beta_2, np.float32),
        "optimizer/decay": np.asarray(decay, np.float32),
    }
    for name, a in m.items():
        out[name + SLOT_M] = np.asarray(a, np.float32)
    for name, a in v.items():
        out[name + SLOT_V] = np.asarray(a, np.float32)
    return out


def optimizer_tensors(iteration: int, learning_rate: float, momentum: float, slots: Dict[str, np.ndarray],
                      decay: float = 0.0) -> Dict[str, np.ndarray]:
    """The same keys as a dict for `write_bundle` (merged with the model variables)."""
    out: Dict[str, np.ndarray] = {
        "optimizer/iter": np.asarray(iteration, np.int64),
        "optimizer/learning_rate": np.asarray(learning_rate, np.float32),
        "optimizer/momentum": np.asarray(momentum, np.float32),
        "optimizer/decay": np.asarray(decay, np.float32),
    }
    for name, v in slots.items():
        out[name + SLOT_INFIX] = np.asarray(v, np.float32)
    return out


# ------------------------------------------------------------------------------ Keras object graph
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
_SLOT_MARK = "/.OPTIMIZER_SLOT/"


def _pb_str(field: int, b: bytes) -> bytes:
    return _put_varint(field << 3 | 2) + _put_varint(len(b)) + b


def _pb_int(field: int, v: int) -> bytes:
    return _put_varint(field << 3 | 0) + _put_varint(v)


def object_graph_proto(checkpoint_keys: Iterable[str]) -> bytes:
    """Serialized `TrackableObjectGraph` (tensorflow/core/protobuf/trackable_object_graph.proto) for
    the variables named by `checkpoint_keys` (full keys, ending in /.ATTRIBUTES/VARIABLE_VALUE).

    Keras' object-based `load_weights` (train.py:137, eval.py:81) does not look tensors up by name: it
    walks this graph from node 0 (the model), matching each object's dependencies by `local_name`
    (`conv1`, `stages` -> `0` -> `stage` -> `layer_with_weights-0` -> `bottleneck` -> `a` -> `kernel`,
    ...), and reads a matched variable from the `checkpoint_key` of its VARIABLE_VALUE attribute.  The
    checkpoint keys ARE those dependency paths, so the graph is the trie of the keys; slot variables
    (`<variable path>/.OPTIMIZER_SLOT/optimizer/<slot>`) hang off the `optimizer` node as
    `slot_variables {original_variable_node_id, slot_name, slot_variable_node_id}`.  Aliases Keras also
    records (`layer-N`, `layer_with_weights-N` on the model, `keras_api`) are optional for restore and
    are not emitted."""
    children: List[Dict[str, int]] = [{}]             # node id -> {local_name: child id}
    attr: Dict[int, Tuple[str, str]] = {}             # variable node id -> (full_name, checkpoint_key)
    slots: List[Tuple[str, str, str]] = []            # (variable path, optimizer path, slot name)

    def node_for(path: str) -> int:
        cur = 0
        for part in path.split("/"):
            nxt = children[cur].get(part)
            if nxt is None:
                nxt = len(children)
                children.append({})
                children[cur][part] = nxt
            cur = nxt
        return cur

    keys = sorted(k for k in checkpoint_keys if k.endswith(VAR_SUFFIX))
    for k in keys:
        path = k[:-len(VAR_SUFFIX)]
        if _SLOT_MARK in path:
            var_path, rest = path.split(_SLOT_MARK, 1)
            opt_path, slot_name = rest.rsplit("/", 1)
            slots.append((var_path, opt_path, slot_name))
        else:
            attr[node_for(path)] = (path, k)
    slot_refs: Dict[int, List[Tuple[int, str, int]]] = {}
    for var_path, opt_path, slot_name in slots:
        var_id, opt_id = node_for(var_path), node_for(opt_path)
        sid = len(children)
        children.append({})
        key = f"{var_path}{_SLOT_MARK}{opt_path}/{slot_name}{VAR_SUFFIX}"
        attr[sid] = (f"{var_path}/{slot_name}", key)
        slot_refs.setdefault(opt_id, []).append((var_id, slot_name, sid))
    out = bytearray()
    for nid, ch in enumerate(children):
        body = bytearray()
        for name, cid in ch.items():
            body += _pb_str(1, _pb_int(1, cid) + _pb_str(2, name.encode("utf-8")))
        if nid in attr:
            full, key = attr[nid]
            body += _pb_str(2, _pb_str(1, b"VARIABLE_VALUE") + _pb_str(2, full.encode("utf-8")) +
                            _pb_str(3, key.encode("utf-8")))
        for var_id, slot_name, sid in slot_refs.get(nid, []):
            body += _pb_str(3, _pb_int(1, var_id) + _pb_str(2, slot_name.encode("utf-8")) + _pb_int(3, sid))
        out += _pb_str(1, bytes(body))
    return bytes(out)


def parse_object_graph(proto: bytes) -> List[dict]:
    """Inverse of `object_graph_proto` (tests; also reads the graph of a real Keras checkpoint):
    [{"children": {local_name: id}, "attributes": [(name, full_name, checkpoint_key)],
      "slot_variables": [(original_variable_node_id, slot_name, slot_variable_node_id)]}, ...]."""
    nodes = []
    for f, _, node in _pb_fields(proto):
        if f != 1:
            continue
        d = {"children": {}, "attributes": [], "slot_variables": []}
        for g, _, v in _pb_fields(node):
            sub = {h: w for h, _, w in _pb_fields(v)}
            if g == 1:
                d["children"][bytes(sub.get(2, b"")).decode("utf-8")] = int(sub.get(1, 0))
            elif g == 2:
                d["attributes"].append(tuple(bytes(sub.get(i, b"")).decode("utf-8") for i in (1, 2, 3)))
            elif g == 3:
                d["slot_variables"].append((int(sub.get(1, 0)), bytes(sub.get(2, b"")).decode("utf-8"),
                                            int(sub.get(3, 0))))
        nodes.append(d)
    return nodes


def _string_tensor(strings: List[bytes]) -> Tuple[bytes, int]:
    """On-disk form of a DT_STRING tensor (tensor_bundle.cc WriteStringTensor): [varint64 length]*,
    masked CRC-32C of the lengths (each as a little-endian uint32), the bytes.  Returns (raw, crc):
    crc covers the lengths, the length checksum and the bytes."""
    raw, crc = bytearray(), 0
    for b in strings:
        raw += _put_varint(len(b))
        crc = crc32c(struct.pack("<I", len(b)), crc)
    cks = struct.pack("<I", mask_crc(crc))
    raw += cks
    crc = crc32c(cks, crc)
    for b in strings:
        raw += b
        crc = crc32c(b, crc)
    return bytes(raw), crc


def _parse_string_tensor(raw: bytes, n: int) -> List[bytes]:
    pos, lens, crc = 0, [], 0
    for _ in range(n):
        v, pos = _get_varint(raw, pos)
        lens.append(v)
        crc = crc32c(struct.pack("<I", v), crc)
    if struct.unpack("<I", raw[pos:pos + 4])[0] != mask_crc(crc):
        raise BundleError("string tensor: length checksum mismatch")
    pos += 4
    out = []
    for v in lens:
        out.append(bytes(raw[pos:pos + v]))
        pos += v
    return out


# ------------------------------------------------------------------------------ writer
class _BlockBuilder:
    def __init__(self, restart_interval: int = 16):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""
        self.interval = restart_interval

    def add(self, key: bytes, value: bytes) -> None:
        shared = 0
        if self.count < self.interval:
            m = min(len(self.last), len(key))
            while shared < m and self.last[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self) -> bytes:
        out = bytes(self.buf)
        out += b"".join(struct.pack("<I", r) for r in self.restarts)
        return out + struct.pack("<I", len(self.restarts))


def _emit_block(f, contents: bytes) -> Tuple[int, int]:
    off = f.tell()
    f.write(contents)
    f.write(b"\x00" + struct.pack("<I", mask_crc(crc32c(contents + b"\x00"))))
    return off, len(contents)


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], block_size: int = 256 << 10,
                 add_suffix: bool = True, state_file: bool = True, object_graph: bool = True) -> None:
    """Write `{prefix}.index` + `{prefix}.data-00000-of-00001` holding `tensors`
    (float32 / int64 / ...), keys sorted like TF's writer does, plus (by default) the
    `_CHECKPOINTABLE_OBJECT_GRAPH` string tensor Keras' object-based `load_weights` walks
    (`object_graph_proto`), so that the reference's `train.py:137` / `eval.py:81` can restore what
    `X3D.save_weights` / `X3DTrainer.save_checkpoint` write."""
    items = sorted(((k + VAR_SUFFIX if add_suffix else k), v) for k, v in tensors.items())
    if object_graph and add_suffix:
        items.append((OBJECT_GRAPH_KEY, object_graph_proto(k for k, _ in items)))
        items.sort(key=lambda kv: kv[0])
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    entries: List[Tuple[bytes, bytes]] = []
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"        # num_shards=1, version{producer=1}
    entries.append((b"", header))
    with open(prefix + ".data-00000-of-00001", "wb") as df:
        for k, arr in items:
            if isinstance(arr, (bytes, bytearray)):                  # scalar DT_STRING tensor
                raw, c = _string_tensor([bytes(arr)])
                e = BundleEntry(dtype=DT_STRING, shape=(), shard_id=0, offset=df.tell(), size=len(raw),
                                crc32c=mask_crc(c))
                df.write(raw)
                entries.append((k.encode("utf-8"), e.serialize()))
                continue
            a = np.asarray(arr, order="C")
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if np.dtype(dt) not in _DT_OF_NP:
                raise BundleError(f"{k}: dtype {a.dtype} not supported")
            raw = a.astype(dt, copy=False).tobytes()
            e = BundleEntry(dtype=_DT_OF_NP[np.dtype(dt)], shape=tuple(a.shape), shard_id=0,
                            offset=df.tell(), size=len(raw), crc32c=mask_crc(crc32c(raw)))
            df.write(raw)
            entries.append((k.encode("utf-8"), e.serialize()))
    with open(prefix + ".index", "wb") as f:
        index = _BlockBuilder(restart_interval=1)
        blk = _BlockBuilder()
        for i, (k, v) in enumerate(entries):
            blk.add(k, v)
            if len(blk.buf) >= block_size or i == len(entries) - 1:
                off, sz = _emit_block(f, blk.finish())
                index.add(k, _put_varint(off) + _put_varint(sz))
                blk = _BlockBuilder()
        moff, msz = _emit_block(f, _BlockBuilder().finish())
        ioff, isz = _emit_block(f, index.finish())
        footer = _put_varint(moff) + _put_varint(msz) + _put_varint(ioff) + _put_varint(isz)
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    if state_file:
        base = os.path.basename(prefix)
        with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
            f.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')
