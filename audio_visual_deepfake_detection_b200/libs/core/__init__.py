from .config import load_default_config, load_config, load_config_for

__all__ = ["load_default_config", "load_config", "load_config_for"]
