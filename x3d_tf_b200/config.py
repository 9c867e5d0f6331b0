"""Configuration for the X3D hot path.

Mirrors the key tree the reference reads through yacs (`configs/default.py:3-140`,
`configs/kinetics/X3D_*.yaml`): `cfg.NETWORK.*`, `cfg.NETWORK.BN.*`, `cfg.DATA.*`,
`cfg.TRAIN.*`, `cfg.TEST.*`.  yacs is not available in this image, so `CfgNode` here is a
small attribute-dict with the three methods the reference's callers use
(`merge_from_file`, `freeze`, `clone`; `train.py:39-41`, `eval.py:28-30`).  A reference YAML
file can be merged unchanged with `cfg.merge_from_file(path)`.
"""
from __future__ import annotations

import copy
from typing import Any, Dict

import yaml


class CfgNode(dict):
    """Attribute-style nested dict (subset of yacs.config.CfgNode)."""

    _FROZEN = "__frozen__"

    def __init__(self, init: Dict[str, Any] | None = None):
        super().__init__()
        object.__setattr__(self, CfgNode._FROZEN, False)
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name: str, value: Any) -> None:
        if object.__getattribute__(self, CfgNode._FROZEN):
            raise AttributeError(f"attempted to set {name} on a frozen CfgNode")
        self[name] = value

    def freeze(self) -> None:
        object.__setattr__(self, CfgNode._FROZEN, True)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self) -> None:
        object.__setattr__(self, CfgNode._FROZEN, False)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def is_frozen(self) -> bool:
        return object.__getattribute__(self, CfgNode._FROZEN)

    def clone(self) -> "CfgNode":
        return CfgNode(copy.deepcopy(_to_plain(self)))

    def merge_from_dict(self, other: Dict[str, Any]) -> None:
        _merge(self, other, path="")

    def merge_from_file(self, path: str) -> None:
        with open(path, "r") as f:
            self.merge_from_dict(yaml.safe_load(f) or {})

    def merge_from_list(self, kv: list) -> None:
        assert len(kv) % 2 == 0
        for k, v in zip(kv[0::2], kv[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f"Non-existent config key: {k}")
            node[parts[-1]] = v


def _to_plain(node: Any) -> Any:
    if isinstance(node, dict):
        return {k: _to_plain(v) for k, v in node.items()}
    return node


def _coerce(old: Any, new: Any, path: str) -> Any:
    # yacs allows int<->float and list<->tuple replacement; YAML reads `5e-5` as a string.
    if isinstance(old, float) and isinstance(new, (int, str)):
        return float(new)
    if isinstance(old, bool) or old is None:
        return new
    if isinstance(old, int) and isinstance(new, float) and float(new).is_integer():
        return int(new)
    if isinstance(old, (list, tuple)) and isinstance(new, (list, tuple)):
        return list(new)
    if type(old) is not type(new):
        raise ValueError(f"type mismatch for config key {path}: {type(old)} vs {type(new)}")
    return new


def _merge(dst: CfgNode, src: Dict[str, Any], path: str) -> None:
    if dst.is_frozen():
        raise AttributeError("cannot merge into a frozen CfgNode")
    for k, v in src.items():
        full = f"{path}.{k}" if path else k
        if k not in dst:
            raise KeyError(f"Non-existent config key: {full}")
        if isinstance(dst[k], CfgNode):
            if not isinstance(v, dict):
                raise ValueError(f"config key {full} is a node")
            _merge(dst[k], v, full)
        else:
            dst[k] = _coerce(dst[k], v, full)


# Defaults: configs/default.py:8-137 of the reference.
_DEFAULTS: Dict[str, Any] = {
    "NETWORK": {
        "C1_TEMP_FILTER": 5,
        "C1_CHANNELS": 12,
        "SCALE_RES2": False,
        "WIDTH_FACTOR": 1.0,
        "DEPTH_FACTOR": 1.0,
        "BOTTLENECK_WIDTH_FACTOR": 1.0,
        "NUM_CLASSES": 400,
        "DROPOUT_RATE": 0.0,
        "WEIGHT_DECAY": 0.00005,
        "BN": {"MOMENTUM": 0.9, "EPS": 1e-5},
    },
    "DATA": {
        "FRAME_RATE": 1,
        "TEMP_DURATION": 1,
        "NUM_INPUT_CHANNELS": 3,
        "TRAIN_JITTER_SCALES": [182, 228],
        "TRAIN_CROP_SIZE": 112,
        "TEST_CROP_SIZE": 160,
        "MEAN": [0.45, 0.45, 0.45],
        "STD": [0.225, 0.225, 0.225],
    },
    "TRAIN": {
        "DATASET_SIZE": 0,
        "BATCH_SIZE": 1,
        "EPOCHS": 1,
        "OPTIMIZER": "SGD",
        "MOMENTUM": 0.9,
        "BASE_LR": 0.1,
        "WARMUP_EPOCHS": 1,
        "WARMUP_LR": 0.01,
    },
    "TEST": {"NUM_SPATIAL_CROPS": 3, "NUM_TEMPORAL_VIEWS": 1, "BATCH_SIZE": 1},
    "WANDB": {
        "ENABLE": False,
        "PROJECT_NAME": "X3D-tf",
        "GROUP_NAME": " ",
        "MODE": "online",
        "TENSORBOARD": True,
    },
}

_KINETICS_COMMON = {
    "NETWORK": {
        "BOTTLENECK_WIDTH_FACTOR": 2.25,
        "C1_CHANNELS": 12,
        "NUM_CLASSES": 400,
        "DROPOUT_RATE": 0.5,
        "WEIGHT_DECAY": 5e-5,
        "BN": {"MOMENTUM": 0.9, "EPS": 1e-5},
    },
    "DATA": {
        "NUM_INPUT_CHANNELS": 3,
        "MEAN": [0.433, 0.404, 0.377],
        "STD": [0.151, 0.148, 0.157],
    },
    "TRAIN": {"EPOCHS": 256, "OPTIMIZER": "sgd", "WARMUP_EPOCHS": 35, "WARMUP_LR": 0.01,
              "MOMENTUM": 0.9},
}

# Per-variant overrides: configs/kinetics/X3D_{XS,S,M,L,XL}.yaml of the reference.
# "TEST3_CROP_SIZE" is the commented 3-crop test size of each YAML (line 20-22); it is kept
# outside the cfg tree (see `three_crop_size`).
_VARIANTS: Dict[str, Dict[str, Any]] = {
    "X3D_XS": {
        "NETWORK": {"WIDTH_FACTOR": 1.0, "DEPTH_FACTOR": 2.2},
        "DATA": {"FRAME_RATE": 12, "TEMP_DURATION": 4, "TRAIN_JITTER_SCALES": [182, 228],
                 "TRAIN_CROP_SIZE": 160, "TEST_CROP_SIZE": 160},
        "TRAIN": {"DATASET_SIZE": 234619, "BATCH_SIZE": 128, "BASE_LR": 0.2},
        "TEST": {"NUM_SPATIAL_CROPS": 1, "NUM_TEMPORAL_VIEWS": 10, "BATCH_SIZE": 12},
    },
    "X3D_S": {
        "NETWORK": {"WIDTH_FACTOR": 1.0, "DEPTH_FACTOR": 2.2},
        "DATA": {"FRAME_RATE": 6, "TEMP_DURATION": 13, "TRAIN_JITTER_SCALES": [182, 228],
                 "TRAIN_CROP_SIZE": 160, "TEST_CROP_SIZE": 160},
        "TRAIN": {"DATASET_SIZE": 234619, "BATCH_SIZE": 64, "BASE_LR": 0.1},
        "TEST": {"NUM_SPATIAL_CROPS": 1, "NUM_TEMPORAL_VIEWS": 10, "BATCH_SIZE": 8},
    },
    "X3D_M": {
        "NETWORK": {"WIDTH_FACTOR": 1.0, "DEPTH_FACTOR": 2.2},
        "DATA": {"FRAME_RATE": 5, "TEMP_DURATION": 16, "TRAIN_JITTER_SCALES": [256, 320],
                 "TRAIN_CROP_SIZE": 224, "TEST_CROP_SIZE": 224},
        "TRAIN": {"DATASET_SIZE": 234584, "BATCH_SIZE": 32, "BASE_LR": 0.05},
        "TEST": {"NUM_SPATIAL_CROPS": 1, "NUM_TEMPORAL_VIEWS": 10, "BATCH_SIZE": 4},
    },
    "X3D_L": {
        "NETWORK": {"WIDTH_FACTOR": 1.0, "DEPTH_FACTOR": 5.0},
        "DATA": {"FRAME_RATE": 5, "TEMP_DURATION": 16, "TRAIN_JITTER_SCALES": [356, 446],
                 "TRAIN_CROP_SIZE": 312, "TEST_CROP_SIZE": 312},
        "TRAIN": {"DATASET_SIZE": 234584, "BATCH_SIZE": 16, "BASE_LR": 0.025},
        "TEST": {"NUM_SPATIAL_CROPS": 1, "NUM_TEMPORAL_VIEWS": 3, "BATCH_SIZE": 8},
    },
    "X3D_XL": {
        "NETWORK": {"WIDTH_FACTOR": 2.9, "DEPTH_FACTOR": 5.0, "SCALE_RES2": True},
        "DATA": {"FRAME_RATE": 5, "TEMP_DURATION": 16, "TRAIN_JITTER_SCALES": [356, 446],
                 "TRAIN_CROP_SIZE": 312, "TEST_CROP_SIZE": 312},
        "TRAIN": {"DATASET_SIZE": 234584, "BATCH_SIZE": 16, "BASE_LR": 0.025},
        "TEST": {"NUM_SPATIAL_CROPS": 1, "NUM_TEMPORAL_VIEWS": 3, "BATCH_SIZE": 8},
    },
}

_THREE_CROP = {"X3D_XS": 182, "X3D_S": 182, "X3D_M": 256, "X3D_L": 356, "X3D_XL": 356}


def get_default_config() -> CfgNode:
    """Same role as `configs/default.py:139-140`."""
    return CfgNode(copy.deepcopy(_DEFAULTS))


def get_config(variant: str, freeze: bool = True) -> CfgNode:
    """Config of a named variant ("X3D_XS" … "X3D_XL"; "-" accepted for "_")."""
    key = variant.upper().replace("-", "_")
    if key not in _VARIANTS:
        raise KeyError(f"unknown X3D variant {variant!r}; known: {sorted(_VARIANTS)}")
    cfg = get_default_config()
    cfg.merge_from_dict(copy.deepcopy(_KINETICS_COMMON))
    cfg.merge_from_dict(copy.deepcopy(_VARIANTS[key]))
    cfg.WANDB.GROUP_NAME = key.replace("_", "-")
    if freeze:
        cfg.freeze()
    return cfg


def three_crop_size(variant: str) -> int:
    """The spatial size each YAML names for 3-crop testing (e.g. `X3D_M.yaml:21`)."""
    return _THREE_CROP[variant.upper().replace("-", "_")]


def variants() -> list:
    return list(_VARIANTS)
