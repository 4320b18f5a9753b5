"""Shim: attribute-access dict standing in for omegaconf.DictConfig."""


class DictConfig(dict):
    def __init__(self, content=None, **kw):
        super().__init__()
        for k, v in dict(content or {}, **kw).items():
            self[k] = DictConfig(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class OmegaConf:
    @staticmethod
    def create(d):
        return DictConfig(d)
