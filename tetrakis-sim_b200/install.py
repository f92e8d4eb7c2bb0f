"""Rebind the reference's entry points onto the B200 path.

The reference has no solver plugin kind (``tetrakis_sim/registry.py`` is an unused scaffold), and
its callers bind ``run_wave_sim`` / ``run_fft`` by name at import time
(``batch.py:13``, ``cli.py:9``, ``evals/dataset.py:10``, ``scripts/bench_wave.py:12``).  So the
drop-in is a rebinding of those names in every already-imported ``tetrakis_sim`` module:

    import tetrakis_sim_b200 as tk
    tk.install()            # tetrakis-batch, tetrakis-eval generate, bench_wave.py now run on the GPU
"""

from __future__ import annotations

import importlib
import sys

_TARGETS = ("tetrakis_sim.physics", "tetrakis_sim.batch", "tetrakis_sim.cli",
            "tetrakis_sim.evals.dataset", "tetrakis_sim")  # fmt: skip
_NAMES = ("run_wave_sim", "run_fft", "SimulationHistory")
_saved: dict = {}


def install(extra_modules=(), structured: bool | None = None) -> list[str]:
    """Replace run_wave_sim / run_fft / SimulationHistory in the reference's modules (and in any
    module named in ``extra_modules``, e.g. a script that did ``from tetrakis_sim.physics import
    run_wave_sim``).  Returns the list of patched ``module.name`` strings.

    ``structured=True`` additionally lets ``path="auto"`` recognise the tetrakis lattices the reference's
    front-ends build (``batch.py:87-114``) and run them on the structured kernel -- 24 bytes per node update
    instead of 44 + 4 * degree, results within 1e-12 of the reference instead of bit-identical; ``None`` leaves
    the choice to ``TETRAKIS_B200_AUTO_STRUCTURED``."""
    from . import physics as ours

    if structured is not None:
        ours.set_auto_structured(structured)

    patched = []
    for modname in (*_TARGETS, *extra_modules):
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
        for name in _NAMES:
            if hasattr(mod, name):
                _saved.setdefault((modname, name), getattr(mod, name))
                setattr(mod, name, getattr(ours, name))
                patched.append(f"{modname}.{name}")
    return patched


def uninstall() -> None:
    from . import physics as ours

    ours.set_auto_structured(None)
    for (modname, name), obj in list(_saved.items()):
        mod = sys.modules.get(modname)
        if mod is not None:
            setattr(mod, name, obj)
        del _saved[(modname, name)]
