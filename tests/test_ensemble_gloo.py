"""CPU, world_size 2 and 3 over gloo: the host side of the sharded ensembles (BASELINE configs[4]; SURVEY.md 8e:
"independent units, sims block-distributed over GPUs") -- per-mask arrays, kick rule, per-simulation CFL clamp, the
contiguous block every rank takes, the gather -- against the literal restatement of ``run_wave_sim`` + ``run_fft`` on the
graph the reference would build for every member.  The arithmetic of a member is the oracle's (tests/
ensemble_numpy_backend.py); on the GPU the same descriptor drives ``ensemble_kernel`` (tests/test_gpu_ensemble.py)."""

import os
import pickle
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import tetrakis_sim_b200 as tk
from oracle import lattice_oracle as lo
from oracle import wave_oracle as wo

RADII, DAMPINGS, DTS = [0.0, 1.5, 2.5], [0.0, 0.03], [0.05, 0.3]  # dt 0.3 is clamped where the hole leaves degree >= 12
SIZE, LAYERS, STEPS = 5, 3, 12


def _grid():
    return tk.sweep_grid(SIZE, LAYERS, RADII, DAMPINGS, DTS)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.ensemble_numpy_backend import FakeEnsembleLib
        from tetrakis_sim_b200 import _lib

        fake = FakeEnsembleLib()
        _lib.lib = lambda: fake
        specs, damp, dt = _grid()
        res, (lo_, hi_) = tk.run_ensemble_sharded(specs, STEPS, c=1.0, dt=dt, damping=damp, dist=dist, device=rank)
        mine, (lo2, hi2) = tk.run_ensemble_sharded(specs, STEPS, c=1.0, dt=dt, damping=damp, dist=dist, device=rank,
                                                    gather=False)
        assert (lo_, hi_) == (lo2, hi2) and fake.calls == [(rank, hi_ - lo_)] * 2
        assert np.array_equal(mine.series, res.series[lo_:hi_])
        with open(os.path.join(out, f"ens{rank}.pkl"), "wb") as f:
            pickle.dump({"lo": lo_, "hi": hi_, "series": res.series, "spectrum": res.spectrum, "bin": res.dominant_bin,
                         "freq": res.dominant_freq, "dt": res.effective_dt, "adj": res.stability_adjusted,
                         "maxdeg": res.max_degree, "kicks": res.kick_nodes, "updates": res.node_updates}, f)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ensemble_matches_the_literal_restatement_member_by_member(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [pickle.load(open(tmp_path / f"ens{r}.pkl", "rb")) for r in range(world)]
    specs, damp, dt = _grid()
    n = len(specs)
    assert n == 12
    # contiguous blocks that partition the members, sizes differing by at most one
    blocks = [(g["lo"], g["hi"]) for g in got]
    assert blocks[0][0] == 0 and blocks[-1][1] == n and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    assert max(h - l for l, h in blocks) - min(h - l for l, h in blocks) <= 1
    for g in got[1:]:  # every rank holds the same gathered result
        for k in ("series", "spectrum", "bin", "freq", "dt", "adj", "maxdeg"):
            assert np.array_equal(g[k], got[0][k]), k
        assert g["kicks"] == got[0]["kicks"] and g["updates"] == got[0]["updates"]
    g = got[0]
    updates = 0
    for i, spec in enumerate(specs):
        G, removed, center = lo.build_case(SIZE, 3, LAYERS, "blackhole", radius=spec.radius)
        kick = lo.pick_kick_node(G, center, 3)
        assert g["kicks"][i] == kick
        nodes, rowptr, colidx, inv_mass, potential = lo.graph_arrays(G)
        deg = np.array([G.degree[v] for v in nodes], dtype=np.float64)
        dt_eff, meta, _ = wo.cfl_clamp(int(deg.max()), 1.0, float(dt[i]), STEPS, float(damp[i]))
        assert g["maxdeg"][i] == int(deg.max()) and g["dt"][i] == dt_eff
        assert bool(g["adj"][i]) == bool(meta["stability_adjusted"])
        u0 = np.zeros(len(nodes))
        u0[nodes.index(kick)] = 1.0
        ref = wo.run_graph(rowptr, colidx, deg, inv_mass, potential, u0, STEPS, dt_eff ** 2, float(damp[i]))["history"]
        series = ref[:, nodes.index(kick)]
        assert np.max(np.abs(g["series"][i] - series)) <= 1e-13 * np.max(np.abs(series))
        freq, spectrum, _ = wo.run_fft(series, dt_eff)
        assert np.max(np.abs(g["spectrum"][i] - spectrum)) <= 1e-12 * np.max(spectrum)
        assert abs(g["freq"][i]) == pytest.approx(abs(freq[wo.folded_bin(spectrum)]), rel=1e-12)
        updates += len(nodes) * STEPS
    assert g["updates"] == updates
    assert g["adj"].any() and not g["adj"].all()
