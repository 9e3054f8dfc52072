"""Import alias for the ``dep-gan-im_b200/`` package directory (a hyphenated name is not importable).

``import depgan_b200`` resolves sub-modules from ``../dep-gan-im_b200/``.
"""
from pathlib import Path as _Path

__path__.append(str(_Path(__file__).resolve().parent.parent / "dep-gan-im_b200"))

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
