"""Attribute-style configuration tree.

Mirrors the three behaviours of the reference's CfgNode that the Stage-I scripts exercise
(ref: nerf/cfgnode.py:36-119 constructor + `__getattr__`, :167-187 `dump`; call sites
eval_stage_rays.py:265-267, train_stage_rays_auto.py:231): build from a nested dict loaded from YAML,
read keys as attributes (missing key -> AttributeError, so `hasattr(cfg.models, "fine")` works,
ref: nerf/train_utils.py:238), and dump back to YAML text.  Merge/freeze/deprecation machinery of the
reference is never called by the hot path and is intentionally not provided.
"""
import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        for key, value in (init_dict or {}).items():
            self[key] = CfgNode(value) if isinstance(value, dict) and not isinstance(value, CfgNode) else value

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def dump(self, **kwargs):
        return yaml.safe_dump(self.to_dict(), **kwargs)

    def clone(self):
        return CfgNode(self.to_dict())

    @classmethod
    def load_yaml(cls, path):
        with open(path, "r") as f:
            return cls(yaml.load(f, Loader=yaml.FullLoader))
