"""Minimal attribute-access config node.

The reference's ``general_utils/cfgnode.py`` (a yacs derivative, 507 lines) is out of scope of
the hot path and is reused as is when the reference tree is on the
path; this stand-in offers the subset the hot path touches (attribute and item access on nested
mappings, mutation after construction) so the package also works standalone.
"""
import copy

import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return CfgNode(copy.deepcopy(dict(self), memo))

    def dump(self, **kwargs):
        def plain(n):
            return {k: plain(v) if isinstance(v, dict) else v for k, v in n.items()}
        return yaml.safe_dump(plain(self), **kwargs)

    @classmethod
    def load_yaml(cls, path):
        with open(path) as f:
            return cls(yaml.safe_load(f))
