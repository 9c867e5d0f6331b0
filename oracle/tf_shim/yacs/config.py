"""Minimal `yacs.config.CfgNode` (attribute dict, merge_from_file, freeze, clone) for
configs/default.py:1-141 and the YAMLs in configs/kinetics/.  TEST INFRASTRUCTURE."""
import ast
import copy

import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__["_frozen"]:
            raise AttributeError("Attempted to set {} to {}, but CfgNode is immutable".format(name, value))
        self[name] = value

    def _set_frozen(self, flag):
        self.__dict__["_frozen"] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def is_frozen(self):
        return self.__dict__["_frozen"]

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    def merge_from_file(self, cfg_filename):
        with open(cfg_filename, "r") as f:
            _merge(yaml.safe_load(f) or {}, self, [])

    def merge_from_list(self, cfg_list):
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            d = self
            parts = full_key.split(".")
            for p in parts[:-1]:
                d = d[p]
            d[parts[-1]] = _coerce(_decode(v), d[parts[-1]], full_key)


def _decode(v):
    if isinstance(v, str):
        try:
            return ast.literal_eval(v)            # yacs does the same: "5e-5" -> 5e-05
        except (ValueError, SyntaxError):
            return v
    return v


def _coerce(new, old, key):
    if type(new) is type(old) or old is None:
        return new
    for a, b in ((list, tuple), (tuple, list)):
        if isinstance(new, a) and isinstance(old, b):
            return b(new)
    if isinstance(new, int) and isinstance(old, float):
        return float(new)
    raise ValueError("Type mismatch ({} vs. {}) for config key: {}".format(type(old), type(new), key))


def _merge(src, dst, path):
    for k, v in src.items():
        full = ".".join(path + [k])
        if k not in dst:
            raise KeyError("Non-existent config key: {}".format(full))
        if isinstance(v, dict):
            _merge(v, dst[k], path + [k])
        else:
            dict.__setitem__(dst, k, _coerce(_decode(v), dst[k], full))
