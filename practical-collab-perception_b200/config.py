"""Attribute-access config dict: what the reference modules expect of ``model_cfg``
(EasyDict in the reference, pcdet/config.py:1-10; modules use attribute access and ``.get``)."""


class CfgDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def cfg_get(cfg, key, default=None):
    """Reads ``key`` from an EasyDict / dict / namespace-like config."""
    if hasattr(cfg, "get"):
        return cfg.get(key, default)
    return getattr(cfg, key, default)
