"""Attribute-access configuration node for the YAML schema of nerf-pytorch/config/*.yml.

Mirrors the part of the reference's CfgNode (nerf/cfgnode.py:36-187) that the scripts use:
construction from a (nested) dict, attribute access that raises AttributeError for missing keys
(so `hasattr(cfg.models, "fine")` gates the fine network, train_dexnerf_rgb.py:132), item access,
and `.dump()` to YAML.  The freeze / merge / rename machinery (cfgnode.py:189-319) is never called
by any reference script and is not provided."""
import copy

import yaml

_VALID_TYPES = (tuple, list, str, int, float, bool, type(None))


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        init_dict = {} if init_dict is None else init_dict
        key_list = [] if key_list is None else key_list
        super().__init__(self._convert(copy.deepcopy(dict(init_dict)), key_list))

    @classmethod
    def _convert(cls, d, key_list):
        for k, v in list(d.items()):
            if isinstance(v, CfgNode):
                continue
            if isinstance(v, dict):
                d[k] = cls(v, key_list=key_list + [k])
            elif not isinstance(v, _VALID_TYPES):
                raise AssertionError("Key {} with value {} is not a valid type; valid types: {}".format(
                    ".".join(key_list + [str(k)]), type(v), _VALID_TYPES))
        return d

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if isinstance(value, dict) and not isinstance(value, CfgNode):
            value = CfgNode(value)
        self[name] = value

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def dump(self, **kwargs):
        return yaml.safe_dump(self.to_dict(), **kwargs)

    @classmethod
    def load_yaml(cls, path):
        with open(path, "r") as f:
            return cls(yaml.load(f, Loader=yaml.FullLoader))

    def __repr__(self):
        return "{}({})".format(self.__class__.__name__, dict.__repr__(self))
