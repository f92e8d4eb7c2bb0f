"""``tetrakis-batch`` on the B200 path, without networkx.

Same flags, console lines and output artefacts as the reference CLI (``tetrakis_sim/batch.py:17-209``):
``<prefix>_spectrum.csv`` (``np.savetxt`` with header ``freq,spectrum``, batch.py:182-188) and
``<prefix>_metadata.json`` (batch.py:192-208).  The lattice is a :class:`LatticeSpec` instead of a
graph, so ``--size 512 --layers 512`` works (the reference needs O(N*size) edges for it).
Plotting is out of scope: like the reference without matplotlib, no PNG is written.

    python -m tetrakis_sim_b200.batch --dim 3 --size 11 --layers 7 --radius 2.5 --steps 50
"""

from __future__ import annotations

import argparse
import ast
import json
import math
import os
from datetime import datetime

import numpy as np

from .lattice import QUADRANTS, LatticeSpec
from .physics import run_fft, run_wave_sim

_MAX_LISTED_NODES = 200_000  # JSON lists of removed / horizon nodes are omitted beyond this


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Run a batch lattice wave simulation on the GPU "
                                            "(flags of tetrakis-batch).")
    p.add_argument("--size", type=int, default=9)
    p.add_argument("--dim", type=int, default=3, choices=[2, 3])
    p.add_argument("--layers", type=int, default=5)
    p.add_argument("--radius", type=float, default=2.5)
    p.add_argument("--steps", type=int, default=40)
    p.add_argument("--c", type=float, default=1.0)
    p.add_argument("--dt", type=float, default=0.2)
    p.add_argument("--damping", type=float, default=0.0)
    p.add_argument("--defect_type", type=str, default="blackhole",
                   choices=["blackhole", "wedge", "none", "singularity"])
    p.add_argument("--outdir", type=str, default="batch_cli_output")
    p.add_argument("--prefix", type=str, default=None)
    p.add_argument("--kick", type=str, default=None)
    p.add_argument("--sing_mass", type=float, default=1000.0)
    p.add_argument("--sing_potential", type=float, default=0.0)
    p.add_argument("--sing_radius", type=float, default=None)
    p.add_argument("--sing_prune_edges", action="store_true")
    return p


def spec_from_args(args) -> LatticeSpec:
    """batch.py:87-114: lattice + defect at the lattice centre."""
    kw = dict(size=args.size, dim=args.dim, layers=args.layers, defect=args.defect_type)
    if args.defect_type == "blackhole":
        kw["radius"] = args.radius
    elif args.defect_type == "singularity":
        kw.update(radius=args.radius if args.sing_radius is None else args.sing_radius,
                  mass=args.sing_mass, potential=args.sing_potential,
                  prune_edges=args.sing_prune_edges)
    return LatticeSpec(**kw)


def event_horizon(spec: LatticeSpec) -> list:
    """Nodes just outside a black hole (defects.py:138-164 with the pre-defect adjacency, as
    batch.py:116-119 calls it): surviving nodes with radius-1 <= dist < radius+1 that had a removed
    neighbour.  In the pre-defect lattice a node's neighbours are its cell mates (same distance, so
    they survive with it), its row and column mates on the same layer and the cells directly above
    and below; sorted (the reference's order is set-iteration order)."""
    if spec.defect != "blackhole":
        return []
    removed_cells = {tuple(c) for c in spec._cells_within(inclusive=False).tolist()}  # (z, r, c)
    if not removed_cells:
        return []
    rows = {(z, r) for z, r, _ in removed_cells}
    cols = {(z, c) for z, _, c in removed_cells}
    ctr, R = spec.defect_center, float(spec.radius)
    span = int(math.floor(R + 1)) + 1
    out = []
    zs = range(spec.layers) if len(ctr) == 2 else range(max(0, ctr[2] - span), min(spec.layers, ctr[2] + span + 1))
    for z in zs:
        for r in range(max(0, int(ctr[0]) - span), min(spec.size, int(ctr[0]) + span + 1)):
            for c in range(max(0, int(ctr[1]) - span), min(spec.size, int(ctr[1]) + span + 1)):
                if (z, r, c) in removed_cells:
                    continue
                if len(ctr) == 2:
                    d = math.hypot(r - ctr[0], c - ctr[1])
                else:
                    d = math.sqrt((r - ctr[0]) ** 2 + (c - ctr[1]) ** 2 + (z - ctr[2]) ** 2)
                if not (R - 1 <= d < R + 1):
                    continue
                if ((z, r) in rows or (z, c) in cols or (z - 1, r, c) in removed_cells
                        or (z + 1, r, c) in removed_cells):
                    for q in QUADRANTS:
                        out.append((r, c, q) if spec.dim == 2 else (r, c, z, q))
    return sorted(out)


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    os.makedirs(args.outdir, exist_ok=True)
    run_id = datetime.now().strftime("%Y%m%d_%H%M%S")
    prefix = args.prefix or f"size{args.size}_radius{args.radius}_layers{args.layers}_steps{args.steps}"
    print("== Tetrakis-Sim Batch CLI ==")
    print(f"Parameters: {vars(args)}")

    spec = spec_from_args(args)
    if args.kick:
        initial_node = ast.literal_eval(args.kick)
        if not spec.contains(initial_node):
            raise ValueError(f"--kick node {initial_node} is not in the graph after defect application")
    else:
        initial_node = spec.kick_node()
    print("kick node:", initial_node)

    history = run_wave_sim(spec, steps=args.steps, initial_data={initial_node: 1.0}, c=args.c,
                           dt=args.dt, damping=args.damping)
    freq, spectrum, values = run_fft(history, initial_node)
    plot_filename = os.path.join(args.outdir, f"{prefix}_fft_node{initial_node}.png")
    csv_filename = os.path.join(args.outdir, f"{prefix}_spectrum.csv")
    np.savetxt(csv_filename, np.column_stack([freq, spectrum]), delimiter=",", header="freq,spectrum")
    print(f"Saved FFT data to {csv_filename}")

    metadata = vars(args).copy()
    metadata["kick_node"] = initial_node
    metadata["effective_dt"] = history.dt
    metadata["stability_adjusted"] = history.metadata.get("stability_adjusted", False)
    metadata["stability_limit"] = history.metadata.get("stability_limit")
    metadata["removed_node_count"] = spec.removed_count
    if spec.removed_count:
        if spec.removed_count <= _MAX_LISTED_NODES:
            removed = sorted(spec.node(int(i)) for i in spec.specials()[0])
            # the reference lists nodes in graph order: r-major, then c, z, q (lattice.py:48-54)
            metadata["removed_nodes"] = removed
            horizon = event_horizon(spec)
            if horizon:
                metadata["event_horizon_nodes"] = horizon
        else:
            metadata["removed_nodes_omitted"] = True
    metadata["csv_filename"] = csv_filename
    metadata["plot_filename"] = plot_filename
    metadata["run_id"] = run_id
    # batch.py:205 takes argmax over the two-sided spectrum of a real signal, where bins k and N-k are equal up to
    # FFT rounding (numpy's own pocketfft picks -0.784 at BASELINE config 1, scipy.fft +0.784: SURVEY.md 0.5).  The
    # tie is broken towards the lower bin here, so the sign is deterministic: compare |dominant_freq|.
    k = int(np.argmax(spectrum))
    mirror = (len(spectrum) - k) % len(spectrum)
    if mirror < k and abs(spectrum[mirror] - spectrum[k]) <= 1e-9 * abs(spectrum[k]):
        k = mirror
    metadata["dominant_freq"] = float(freq[k])
    metadata["dominant_freq_tie"] = "bins k and N-k of a real signal tie to rounding; the lower bin is reported"
    json_filename = os.path.join(args.outdir, f"{prefix}_metadata.json")
    with open(json_filename, "w") as f:
        json.dump(metadata, f, indent=2)
    print(f"Saved run metadata to {json_filename}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
