"""YAML config loading with the reference's semantics (src/utils/config_utils.py:23-71): a path is tried relative
to the cwd first, then relative to the repository root; a missing/unreadable file or a non-mapping document yields
``{}`` silently (callers supply per-key defaults); results are cached per resolved path."""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

try:
    import yaml
except Exception:  # pragma: no cover - PyYAML is present in the image
    yaml = None

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ConfigManager:
    def __init__(self) -> None:
        self._cache: Dict[str, Dict[str, Any]] = {}

    @staticmethod
    def _resolve(path: str) -> str:
        if os.path.exists(path):
            return os.path.abspath(path)
        alt = os.path.join(_REPO_ROOT, path)
        return os.path.abspath(alt) if os.path.exists(alt) else path

    def load_config(self, path: str, defaults: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        key = self._resolve(path)
        if key not in self._cache:
            cfg: Dict[str, Any] = {}
            if yaml is not None and os.path.isfile(key):
                try:
                    with open(key, "r", encoding="utf-8") as f:
                        doc = yaml.safe_load(f)
                    cfg = doc if isinstance(doc, dict) else {}
                except Exception:
                    cfg = {}
            self._cache[key] = cfg
        cfg = self._cache[key]
        if defaults:
            merged = dict(defaults)
            merged.update(cfg)
            return merged
        return cfg


def load_yaml(path: str, defaults: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    return ConfigManager().load_config(path, defaults=defaults)
