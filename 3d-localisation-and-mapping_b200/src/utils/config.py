"""``ConfigLoader`` stand-in: the reference's loader and ``variables.cfg`` are not shipped
(``/root/reference/task_def.py:16,229``).  The lift only needs ``img_size``, ``depth_width``
and ``depth_height`` (``task_def.py:133-142``); other attributes are read from an INI file
if one is given, so the reference's call shape ``ConfigLoader(path, data_folder)`` works."""
from __future__ import annotations

import configparser
import os


class ConfigLoader:
    DEFAULTS = dict(img_size=640, depth_width=192, depth_height=256, display_3d_pose=False)

    def __init__(self, config_path=None, data_folder="gold_std"):
        self.data_folder = data_folder
        for k, v in self.DEFAULTS.items():
            setattr(self, k, v)
        if config_path and os.path.exists(config_path):
            cp = configparser.ConfigParser()
            cp.read(config_path)
            for section in cp.sections():
                for k, v in cp.items(section):
                    v = v.replace("{data}", data_folder)
                    for cast in (int, float):
                        try:
                            v = cast(v)
                            break
                        except (TypeError, ValueError):
                            continue
                    if isinstance(v, str) and v.lower() in ("true", "false"):
                        v = v.lower() == "true"
                    setattr(self, k, v)
