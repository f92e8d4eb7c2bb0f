#!/usr/bin/env python3
"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

Imports ``tetrakis_sim`` read-only from /root/reference, runs ``run_wave_sim`` / ``run_fft`` on a
fixed case list and writes ``tests/golden/<case>.npz``.  The reference ships no numeric goldens
for u(t); these files are what pins ``oracle/`` and the CUDA path.  Nothing here
is read at test time except the ``.npz`` outputs.
"""

from __future__ import annotations

import json
import os
import sys
import warnings

import networkx as nx
import numpy as np

sys.path.insert(0, "/root/reference")
from tetrakis_sim.defects import apply_defect  # noqa: E402
from tetrakis_sim.lattice import build_sheet  # noqa: E402
from tetrakis_sim.physics import run_fft, run_wave_sim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FULL_HISTORY_MAX_VALUES = 150_000  # (steps+1)*N above this -> store snapshots only


def kick_rule(G, center, dim):
    """batch.py:124-151, executed with the reference's own graph."""
    if dim == 3:
        pool = [n for n in G if len(n) > 3 and n[2] == center[2]]
    else:
        pool = list(G.nodes)
    cand = [n for n in pool if G.degree[n] > 0 and not G.nodes[n].get("singular", False)]
    pool = cand if cand else pool
    return min(pool, key=lambda n: (n[0] - center[0]) ** 2 + (n[1] - center[1]) ** 2)


def encode_nodes(nodes):
    if nodes and isinstance(nodes[0], tuple):
        return np.array([[*n[:-1], "ABCD".index(n[-1])] for n in nodes], dtype=np.int64)
    return np.array(nodes, dtype=np.int64).reshape(-1, 1)


def flatten(G):
    nodes = list(G.nodes)
    index = {n: i for i, n in enumerate(nodes)}
    rowptr = [0]
    col = []
    for n in nodes:
        col.extend(index[m] for m in G[n])
        rowptr.append(len(col))
    deg = [G.degree[n] for n in nodes]
    mass = []
    pot = []
    for n in nodes:
        m = float(G.nodes[n].get("mass", 1.0))
        mass.append(1.0 / (m if m > 0.0 else 1.0))
        pot.append(float(G.nodes[n].get("potential", 0.0)))
    return nodes, index, rowptr, col, deg, mass, pot


def record(name, G, *, steps, c, dt, damping, initial_data, probe, lattice=None, snapshots=None):
    nodes, index, rowptr, col, deg, inv_mass, pot = flatten(G)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        hist = run_wave_sim(G, steps=steps, initial_data=initial_data, c=c, dt=dt, damping=damping)
    warned = [str(w.message) for w in caught if issubclass(w.category, RuntimeWarning)]
    if probe is None:
        probe = next(iter(hist[0].keys()))
    freq, spectrum, values = run_fft(hist, probe)
    n = len(nodes)
    full = (steps + 1) * n <= FULL_HISTORY_MAX_VALUES
    if full:
        snap_steps = np.arange(steps + 1)
    else:
        snap_steps = np.array(sorted(set(snapshots or [0, 1, 2, steps // 2, steps])), dtype=np.int64)
    states = np.array([[hist[t][nd] for nd in nodes] for t in snap_steps], dtype=np.float64)
    u0 = np.array([hist[0][nd] for nd in nodes], dtype=np.float64)
    meta = dict(hist.metadata)
    out = dict(
        nodes=encode_nodes(nodes),
        rowptr=np.array(rowptr, dtype=np.int64),
        colidx=np.array(col, dtype=np.int32),
        deg=np.array(deg, dtype=np.int64),
        inv_mass=np.array(inv_mass, dtype=np.float64),
        potential=np.array(pot, dtype=np.float64),
        u0=u0,
        steps=steps,
        c=c,
        dt=dt,
        damping=damping,
        effective_dt=hist.dt,
        metadata=json.dumps(meta, default=float),
        warning=json.dumps(warned),
        snap_steps=snap_steps,
        states=states,
        probe_index=index[probe],
        probe_values=np.asarray(values, dtype=np.float64),
        fft_freq=freq,
        fft_spectrum=spectrum,
        dominant_freq=float(freq[np.argmax(spectrum)]),
        lattice=json.dumps(lattice or {}),
        extra_history0_keys=len(hist[0]) - n,
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name:34s} N={n:6d} nnz={len(col):8d} steps={steps:5d} full={full} dt_eff={hist.dt:.6f}")


def lattice_case(name, size, dim, layers, defect, *, steps, dt=0.2, c=1.0, damping=0.01,
                 kick="batch", snapshots=None, **dk):  # fmt: skip
    G = build_sheet(size=size, dim=dim, layers=layers)
    center = (size // 2, size // 2) if dim == 2 else (size // 2, size // 2, layers // 2)
    kwargs = {}
    spec = dict(size=size, dim=dim, layers=layers if dim == 3 else 1, defect=defect)
    if defect == "blackhole":
        kwargs = {"center": dk.get("center", center), "radius": dk["radius"]}
    elif defect == "wedge":
        kwargs = {"center": center} if dim == 2 else {"center": center[:2], "layer": center[2]}
    elif defect == "singularity":
        kwargs = {
            "center": center,
            "mass": dk.get("mass", 1000.0),
            "potential": dk.get("potential", 0.0),
            "radius": dk["radius"],
            "prune_edges": dk.get("prune_edges", False),
        }
    spec.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in kwargs.items()})
    G, removed = apply_defect(G, defect_type=defect, return_removed=True, **kwargs)
    spec["removed_node_count"] = len(removed)
    if kick == "batch":
        knode = kick_rule(G, center, dim)
    elif kick == "bench":  # scripts/bench_wave.py:43-46
        knode = (size // 2, size // 2, "A") if dim == 2 else (size // 2, size // 2, layers // 2, "A")
    else:
        knode = kick
    spec["kick"] = [*knode[:-1], "ABCD".index(knode[-1])]
    record(name, G, steps=steps, c=c, dt=dt, damping=damping, initial_data={knode: 1.0},
           probe=knode, lattice=spec, snapshots=snapshots)  # fmt: skip


def main():
    # --- generic graphs (physics.py knows nothing about tetrakis structure) -------------------
    record("path3_clamped", nx.path_graph(3), steps=3, c=1.0, dt=5.0, damping=0.0,
           initial_data=None, probe=None)  # tests/test_physics.py:19-27
    G = nx.Graph()
    G.add_edge(0, 1)
    G.nodes[0]["mass"] = 10.0
    G.nodes[1]["potential"] = 2.0
    record("two_node_mass_potential", G, steps=5, c=1.0, dt=0.1, damping=0.0,
           initial_data={0: 1.0}, probe=0)  # tests/test_physics_mass_potential.py
    rng = np.random.default_rng(7)
    G = nx.gnp_random_graph(160, 0.06, seed=3)
    G.add_edge(5, 5)  # self-loop: degree counts it twice, adjacency once
    G.add_node(1000)  # isolated, non-contiguous label
    for n in list(G.nodes)[::7]:
        G.nodes[n]["mass"] = float(rng.uniform(0.5, 20.0))
    for n in list(G.nodes)[::11]:
        G.nodes[n]["potential"] = float(rng.uniform(0.0, 3.0))
    G.nodes[3]["mass"] = -4.0  # non-positive mass -> 1.0 (physics.py:99-100)
    kicks = {int(n): float(rng.normal()) for n in rng.choice(160, 12, replace=False)}
    record("random_graph_multi_kick", G, steps=60, c=0.8, dt=0.3, damping=0.02,
           initial_data=kicks, probe=17)
    G = nx.grid_2d_graph(9, 7)
    G = nx.convert_node_labels_to_integers(G)
    record("grid9x7_unclamped", G, steps=80, c=1.0, dt=0.3, damping=0.0,
           initial_data={31: 1.0, 2: -0.5}, probe=31)

    # --- tetrakis lattices, every defect (SURVEY.md 8c probe list) ----------------------------
    lattice_case("t2d_s7_none", 7, 2, 1, "none", steps=40)
    lattice_case("t2d_s7_wedge", 7, 2, 1, "wedge", steps=40)
    lattice_case("t2d_s7_blackhole_r2", 7, 2, 1, "blackhole", steps=40, radius=2.0)
    lattice_case("t2d_s7_sing_m50_v25", 7, 2, 1, "singularity", steps=40, radius=1.0,
                 mass=50.0, potential=25.0)
    lattice_case("t2d_s7_sing_m2000_prune", 7, 2, 1, "singularity", steps=40, radius=1.5,
                 mass=2000.0, prune_edges=True)
    lattice_case("t3d_s7l5_none", 7, 3, 5, "none", steps=25)
    lattice_case("t3d_s7l5_wedge", 7, 3, 5, "wedge", steps=25)
    lattice_case("t3d_s7l5_blackhole_r2p5", 7, 3, 5, "blackhole", steps=25, radius=2.5)
    lattice_case("t3d_s7l5_blackhole_cyl", 7, 3, 5, "blackhole", steps=25, radius=1.6,
                 center=(3, 3))
    lattice_case("t3d_s7l5_sing_prune_v3", 7, 3, 5, "singularity", steps=25, radius=1.0,
                 mass=200.0, potential=3.0, prune_edges=True)
    # BASELINE config 1 exactly: tetrakis-batch --dim 3 --size 11 --layers 7 --radius 2.5 --steps 50
    lattice_case("cfg1_s11l7_blackhole_r2p5", 11, 3, 7, "blackhole", steps=50, dt=0.2,
                 damping=0.0, radius=2.5, snapshots=[0, 1, 2, 10, 25, 49, 50])
    # BASELINE config 5 unit: size 13, layers 7, dt 0.2 -> clamped
    lattice_case("cfg5_s13l7_none", 13, 3, 7, "none", steps=50, dt=0.2, damping=0.0,
                 snapshots=[0, 1, 25, 50])
    # README singularity example shape (README.md:201-206): size 13, layers 7, mass 2000, prune
    lattice_case("cfg4small_s13l7_sing_m2000_prune", 13, 3, 7, "singularity", steps=50, dt=0.2,
                 damping=0.0, radius=2.5, mass=2000.0, prune_edges=True, snapshots=[0, 1, 25, 50])
    # bench_wave.py shapes (kick = bench rule, dt 0.1)
    lattice_case("bench2d_s21", 21, 2, 1, "none", steps=50, dt=0.1, damping=0.0, kick="bench",
                 snapshots=[0, 1, 25, 50])
    # long run: rounding-error growth ("Parity at 10k steps")
    lattice_case("t2d_s15_none_2000steps", 15, 2, 1, "none", steps=2000, dt=0.1, damping=0.0,
                 kick="bench", snapshots=[0, 50, 500, 1000, 2000])


def cli_goldens():
    """Artefacts of the reference's own entry points (tetrakis-batch, tetrakis-eval generate)."""
    import contextlib
    import io
    import shutil
    import tempfile

    from tetrakis_sim import batch as ref_batch
    from tetrakis_sim.evals.dataset import generate_defect_classification_jsonl

    tmp = tempfile.mkdtemp()
    try:
        for tag, argv in {
            "batch_cfg1": ["--dim", "3", "--size", "11", "--layers", "7", "--radius", "2.5", "--steps", "50"],
            "batch_sing2d": ["--dim", "2", "--size", "9", "--steps", "30", "--defect_type", "singularity",
                             "--radius", "1.5", "--sing_mass", "200", "--sing_potential", "2.5",
                             "--sing_prune_edges", "--damping", "0.01"],
            "batch_wedge3d": ["--dim", "3", "--size", "7", "--layers", "5", "--steps", "25",
                              "--defect_type", "wedge", "--dt", "0.15"],
        }.items():
            old = sys.argv
            sys.argv = ["tetrakis-batch", *argv, "--outdir", tmp, "--prefix", tag]
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref_batch.main()
            sys.argv = old
            shutil.copy(os.path.join(tmp, f"{tag}_spectrum.csv"), os.path.join(HERE, f"{tag}_spectrum.csv"))
            meta = json.load(open(os.path.join(tmp, f"{tag}_metadata.json")))
            for k in ("run_id", "outdir", "csv_filename", "plot_filename"):
                meta.pop(k, None)
            meta["argv"] = argv
            json.dump(meta, open(os.path.join(HERE, f"{tag}_metadata.json"), "w"), indent=1, sort_keys=True)
            print("batch golden", tag)
        for tag, kw in {
            "evals_2d": dict(n_per_class=3, seed=0, size=9, dim=2, layers=5, steps=30, c=1.0, dt=0.2, damping=0.0),
            "evals_3d": dict(n_per_class=2, seed=5, size=7, dim=3, layers=5, steps=20, c=1.0, dt=0.3, damping=0.01),
        }.items():
            out = os.path.join(HERE, f"{tag}.jsonl")
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                generate_defect_classification_jsonl(out, **kw)
            json.dump(kw, open(os.path.join(HERE, f"{tag}_args.json"), "w"))
            print("evals golden", tag)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cfg5_spotcheck():
    """64 random points (seed 0) of the radius x damping x dt grid of
    65 536 simulations (size 13, layers 7, 50 steps), run through the unmodified reference."""
    radii, dampings, dts = np.linspace(0.5, 4.0, 64), np.linspace(0.0, 0.05, 32), np.linspace(0.02, 0.2, 32)
    rng = np.random.default_rng(0)
    picks = np.sort(rng.choice(64 * 32 * 32, 64, replace=False))
    series, spectra, eff, kicks = [], [], [], []
    for g in picks:
        ir, idamp, idt = g // 1024, (g // 32) % 32, g % 32
        G = build_sheet(size=13, dim=3, layers=7)
        center = (6, 6, 3)
        G, _ = apply_defect(G, defect_type="blackhole", return_removed=True, center=center, radius=float(radii[ir]))
        k = kick_rule(G, center, 3)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            h = run_wave_sim(G, steps=50, initial_data={k: 1.0}, c=1.0, dt=float(dts[idt]), damping=float(dampings[idamp]))
        f, sp, v = run_fft(h, k)
        series.append(v)
        spectra.append(sp)
        eff.append(h.dt)
        kicks.append([*k[:-1], "ABCD".index(k[-1])])
    np.savez_compressed(os.path.join(HERE, "cfg5_spotcheck.npz"), picks=picks, series=np.array(series),
                        spectra=np.array(spectra), effective_dt=np.array(eff), kicks=np.array(kicks))
    print("cfg5 spot-check golden:", len(picks), "grid points")


if __name__ == "__main__":
    if "--cfg5-only" in sys.argv:
        cfg5_spotcheck()
        sys.exit(0)
    if "--cli-only" not in sys.argv:
        main()
        cfg5_spotcheck()
    cli_goldens()
