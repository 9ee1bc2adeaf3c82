"""Hydra-1.0 `compose` restated with PyYAML (hydra/omegaconf are not in the image).

The reference builds its flat `hparams` with
`hydra.experimental.compose(config_name="config", overrides=['model=imitation'])`
(/root/reference/train.py:13,17,95). The semantics needed by the BC block are small:
root file -> `defaults` list of {group: file} -> each file merged at the root when it
starts with `# @package _global_`, else under its group name -> `group=file` overrides
swap a default, `key=value` overrides set a scalar -> `${key}` / `${now:fmt}` interpolation.
The result is a plain dict that also allows attribute access (hparams.pytorch_seed,
train.py:103) and item assignment (hparams['camera'] = camera, train.py:99).
"""
from __future__ import annotations

import datetime
import os
import re
from typing import Iterable, Optional

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")


class HParams(dict):
    """dict with attribute access, like the DictConfig the reference indexes both ways."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _load(path: str):
    with open(path) as f:
        text = f.read()
    is_global = bool(re.search(r"^#\s*@package\s+_global_", text, re.M))
    return (yaml.safe_load(text) or {}), is_global


def _interpolate(cfg: dict) -> None:
    now = datetime.datetime.now()

    def sub(v, depth=0):
        if isinstance(v, str):
            def rep(m):
                key = m.group(1)
                if key.startswith("now:"):
                    return now.strftime(key[4:])
                return str(sub(cfg[key], depth + 1)) if key in cfg and depth < 8 else m.group(0)
            return re.sub(r"\$\{([^}]+)\}", rep, v)
        if isinstance(v, dict):
            return {k: sub(x, depth) for k, x in v.items()}
        if isinstance(v, list):
            return [sub(x, depth) for x in v]
        return v

    for k in list(cfg):
        cfg[k] = sub(cfg[k])


def compose(config_name: str = "config", overrides: Optional[Iterable[str]] = None,
            config_path: str = CONFIG_DIR) -> HParams:
    root, _ = _load(os.path.join(config_path, config_name + ("" if config_name.endswith(".yaml") else ".yaml")))
    defaults = root.pop("defaults", []) or []
    choice = {}
    for d in defaults:
        (group, fname), = d.items()
        choice[group] = fname
    scalars = {}
    for ov in overrides or ():
        key, _, val = ov.partition("=")
        if key in choice or os.path.isdir(os.path.join(config_path, key)):
            choice[key] = val
        else:
            scalars[key] = yaml.safe_load(val)
    out = HParams()
    for group, fname in choice.items():
        if fname in (None, "null"):
            continue
        fname = fname if fname.endswith(".yaml") else fname + ".yaml"
        body, is_global = _load(os.path.join(config_path, group, fname))
        if is_global:
            out.update(body)
        else:
            out[group] = HParams(body)
    out.update(root)          # root keys (log_dir, data_dir) are defined after the defaults list
    out.update(scalars)
    _interpolate(out)
    for g in list(out):
        if isinstance(out[g], dict) and not isinstance(out[g], HParams):
            out[g] = HParams(out[g])
        if isinstance(out[g], HParams):
            _interpolate_group(out[g], out)
    return out


def _interpolate_group(group: HParams, root: HParams) -> None:
    for k, v in list(group.items()):
        if isinstance(v, str):
            group[k] = re.sub(r"\$\{([^}]+)\}", lambda m: str(root.get(m.group(1), m.group(0))), v)
