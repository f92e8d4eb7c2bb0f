"""``tetrakis-eval generate`` as one ensemble launch.

Reference: ``tetrakis_sim/evals/dataset.py:60-165`` builds 4 classes x ``n_per_class`` simulations in a
serial loop (build_sheet -> seeded parameter draw -> apply_defect -> kick rule -> run_wave_sim ->
run_fft -> features -> one JSONL line).  Here the same parameter draws (same ``random.Random(seed)``
call sequence, dataset.py:76,100,110-113) define a list of :class:`LatticeSpec`, all simulations run
in one ``tk_ensemble_run`` call, and the records are assembled with the same keys (schema version 1).

    python -m tetrakis_sim_b200.evals generate --out data.jsonl --n-per-class 10 --size 11 --dim 2
"""

from __future__ import annotations

import argparse
import json
import random
from pathlib import Path

from .ensemble import mask_arrays, run_ensemble
from .lattice import LatticeSpec

SCHEMA_VERSION = 1
DEFECT_TYPES = ("none", "wedge", "blackhole", "singularity")


def draw_specs(n_per_class, seed, size, dim, layers):
    """The reference's parameter draws, in its order.  Returns [(label, spec, extra_params)]."""
    rng = random.Random(seed)
    out = []
    for label in DEFECT_TYPES:
        for _ in range(int(n_per_class)):
            extra = {}
            kw = dict(size=size, dim=dim, layers=layers, defect=label)
            if label == "blackhole":
                r = float(rng.choice([1.5, 2.0, 2.5]))
                kw["radius"] = r
                extra = {"radius": r}
            elif label == "wedge":
                if dim == 3:
                    extra = {"layer": int(layers // 2)}
            elif label == "singularity":
                sr = float(rng.choice([0.0, 0.5, 1.0]))
                mass = float(rng.choice([50.0, 200.0, 1000.0]))
                pot = float(rng.choice([0.0, 25.0]))
                prune = bool(rng.choice([False, True]))
                kw.update(radius=sr, mass=mass, potential=pot, prune_edges=prune)
                extra = {"radius": sr, "mass": mass, "potential": pot, "prune_edges": prune}
            out.append((label, LatticeSpec(**kw), extra))
    return out


def generate_defect_classification_jsonl(out_path, *, n_per_class: int = 10, seed: int = 0,
                                         size: int = 11, dim: int = 2, layers: int = 5,
                                         steps: int = 40, c: float = 1.0, dt: float = 0.2,
                                         damping: float = 0.0, device: int = 0) -> Path:
    out_path = Path(out_path)
    out_path.parent.mkdir(parents=True, exist_ok=True)
    drawn = draw_specs(n_per_class, seed, size, dim, layers)
    specs = [s for _, s, _ in drawn]
    # one launch: simulations, probe spectra and the spectral / time-series features all on the device; only the
    # 18 feature values per simulation come back
    res = run_ensemble(specs, steps, c=c, dt=dt, damping=damping, device=device, features=True, series=False,
                       spectrum=False)
    cache: dict = {}
    with out_path.open("w", encoding="utf-8") as f:
        for i, (label, spec, extra) in enumerate(drawn):
            key = (spec.defect, spec.radius, spec.mass, spec.potential, spec.prune_edges)
            if key not in cache:
                _, exc, _, max_degree, n_alive, deg, alive = mask_arrays(spec, want_degrees=True)
                singular = int(spec.specials()[0].size) if spec.defect == "singularity" else 0
                cache[key] = (max_degree, n_alive, deg, int(deg[alive].sum()) // 2, singular)
            max_degree, n_alive, deg, n_edges, singular = cache[key]
            kick = res.kick_nodes[i]
            ctr = spec.lattice_center
            params = {"size": size, "dim": dim, "layers": layers, "steps": steps, "c": c, "dt_requested": dt, "damping": damping}
            params.update(extra)
            params["dt_used"] = float(res.effective_dt[i])
            feats = {
                "n_nodes": float(n_alive),
                "n_edges": float(n_edges),
                "avg_degree": float((2.0 * n_edges / n_alive) if n_alive > 0 else 0.0),
                "max_degree": float(max_degree),
                "removed_node_count": float(spec.removed_count),
                "singular_node_count": float(singular),
            }
            feats.update(res.feature_dict(i))  # ensemble_kernel epilogue (kernels_ensemble.cuh:ens_features)
            feats["kick_degree"] = float(deg[spec.index(kick)])
            feats["kick_dist2_center"] = float((kick[0] - ctr[0]) ** 2 + (kick[1] - ctr[1]) ** 2)
            rec = {"schema_version": SCHEMA_VERSION, "id": f"dc_{i + 1:05d}",
                   "task": "defect_classification", "label": label, "params": params, "features": feats}
            f.write(json.dumps(rec, sort_keys=True) + "\n")
    return out_path


def main(argv=None) -> int:
    p = argparse.ArgumentParser(prog="tetrakis-eval (B200)")
    sub = p.add_subparsers(dest="cmd", required=True)
    g = sub.add_parser("generate", help="Generate an eval dataset (JSONL)")
    g.add_argument("--out", required=True)
    g.add_argument("--n-per-class", type=int, default=10)
    g.add_argument("--seed", type=int, default=0)
    g.add_argument("--size", type=int, default=11)
    g.add_argument("--dim", type=int, default=2, choices=[2, 3])
    g.add_argument("--layers", type=int, default=5)
    g.add_argument("--steps", type=int, default=40)
    g.add_argument("--c", type=float, default=1.0)
    g.add_argument("--dt", type=float, default=0.2)
    g.add_argument("--damping", type=float, default=0.0)
    a = p.parse_args(argv)
    path = generate_defect_classification_jsonl(a.out, n_per_class=a.n_per_class, seed=a.seed,
                                                size=a.size, dim=a.dim, layers=a.layers,
                                                steps=a.steps, c=a.c, dt=a.dt, damping=a.damping)
    print(f"wrote: {path}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
