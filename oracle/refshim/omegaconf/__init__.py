"""omegaconf stand-in: attribute-access dict config (plumbing)."""
import yaml


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return DictConfig(v) if isinstance(v, dict) and not isinstance(v, DictConfig) else v

    def __setattr__(self, k, v):
        self[k] = v


class OmegaConf:
    @staticmethod
    def create(d=None):
        return DictConfig(d or {})

    @staticmethod
    def load(path):
        with open(path) as f:
            return DictConfig(yaml.safe_load(f))

    @staticmethod
    def to_container(cfg, **k):
        return dict(cfg)

    @staticmethod
    def save(cfg, path):
        with open(path, "w") as f:
            yaml.safe_dump(dict(cfg), f)
