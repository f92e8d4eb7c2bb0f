"""Importable alias of the ``tetrakis-sim_b200/`` package directory.

The product package lives in ``tetrakis-sim_b200/`` (a hyphen cannot appear in an ``import``
statement); this stub points ``tetrakis_sim_b200.__path__`` at that directory so that
``import tetrakis_sim_b200.physics`` resolves to ``tetrakis-sim_b200/physics.py``.
"""

import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tetrakis-sim_b200")
__path__.insert(0, _real)

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
