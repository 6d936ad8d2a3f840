"""YAML -> Namespace exactly like the reference driver (city_sender.py:47-223, function.py:24-32), including the
`--config_mod "a.b=v c.d=v"` override strings (values parsed as Python literals, falling back to str)."""
import argparse
import ast
import os

import yaml

DEFAULT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "mine.yml")


def dict2namespace(config):
    ns = argparse.Namespace()
    for key, value in config.items():
        setattr(ns, key, dict2namespace(value) if isinstance(value, dict) else value)
    return ns


def load_config(path=None, config_mod=(), device="cuda"):
    with open(path or DEFAULT) as f:
        raw = yaml.safe_load(f)
    for item in config_mod:
        for tok in item.split():
            key, _, val = tok.partition("=")
            try:
                val = ast.literal_eval(val)
            except (ValueError, SyntaxError):
                pass
            node = raw
            parts = key.split(".")
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = val
    cfg = dict2namespace(raw)
    cfg.device = device
    return cfg
