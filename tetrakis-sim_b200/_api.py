"""Public surface of the package (re-exported by ``tetrakis_sim_b200``)."""

from ._build import build_library
from ._lib import (BackendUnavailable, NoDeviceError, Plan, TkError, device_count, device_info,
                   device_rates, dft_abs)
from .compiler import compile_graph, compile_lattice, structured_from_graph
from .ensemble import EnsembleResult, run_ensemble, run_ensemble_sharded, sweep_grid
from .install import install, uninstall
from .lattice import LatticeSpec
from .physics import SimulationHistory, dominant_frequency, run_fft, run_fft_many, run_wave_sim
from .sheet import run_sheet_sharded

__all__ = [
    "run_wave_sim", "run_fft", "run_fft_many", "SimulationHistory", "dominant_frequency", "LatticeSpec",
    "Plan", "compile_graph", "compile_lattice", "structured_from_graph", "dft_abs",
    "device_count", "device_info", "device_rates", "install", "uninstall", "build_library",
    "run_ensemble", "run_ensemble_sharded", "sweep_grid", "EnsembleResult", "run_sheet_sharded",
    "BackendUnavailable", "NoDeviceError", "TkError",
]  # fmt: skip
