"""Debug aid: the cross scheme against the oracle, errors split into cross cells / bulk cells."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["TK_FX_MIN_NODES"] = "0"
os.environ["TK_FX_MAXFRAC"] = "95"
import tetrakis_sim_b200 as tk
from oracle import wave_oracle as wo

def oracle(spec, u0, steps, coeff, damping, probes):
    n = spec.num_slots
    flags, im, pot = np.full(n, 3, np.uint8), np.ones(n), np.zeros(n)
    si, sf, sm, sv = spec.specials()
    flags[si], im[si], pot[si] = sf, sm, sv
    return wo.run_structured(spec.size, spec.layers, flags, im, pot, spec.corrections(), u0.copy(), steps, coeff, damping, probes=probes)

def case(spec, K, steps, damping=0.0):
    os.environ["TK_LAT_FUSE"] = str(K)
    kick = spec.kick_node()
    kidx = spec.index(kick)
    plan = tk.compile_lattice(spec)
    coeff = 0.9 / plan.max_degree
    plan.set_state_sparse([kidx], [1.0])
    out = plan.run(steps, coeff, damping, probes=[kidx])
    u, up = plan.get_state(want_prev=True)
    info = plan.cross
    plan.close()
    u0 = np.zeros(spec.num_slots); u0[kidx] = 1.0
    ref = oracle(spec, u0, steps, coeff, damping, [kidx])
    S, L = spec.size, spec.layers
    # cross mask from the special cells + probe
    si = spec.specials()[0]
    cells = np.unique(np.concatenate([si >> 2, [kidx >> 2], np.array([a >> 2 for a, b, s in spec.corrections()] + [b >> 2 for a, b, s in spec.corrections()], dtype=np.int64)])).astype(np.int64)
    rows, cols = np.unique((cells // S) % S), np.unique(cells % S)
    m = np.zeros((L, S, S, 4), bool); m[:, rows] = True; m[:, :, cols] = True
    m = m.reshape(-1)
    den = np.abs(ref["u"]).max()
    e = np.abs(u - ref["u"]); ep = np.abs(up - ref["uprev"])
    print(f"S={S} L={L} {spec.defect:11s} K={K:2d} steps={steps:3d} g={damping}: info={info['steps_per_pass']},{info['cross_rows']}x{info['cross_cols']} "
          f"cross u {e[m].max()/den:.1e} up {ep[m].max()/den:.1e} | bulk u {e[~m].max()/den:.1e} up {ep[~m].max()/den:.1e} | probes {np.abs(out['probes']-ref['probes']).max()/den:.1e}", flush=True)

if __name__ == "__main__":
    for K, steps in [(2, 2), (2, 4), (3, 3), (8, 8), (8, 16)]:
        case(tk.LatticeSpec(size=24, dim=2, defect="blackhole", radius=3.3), K, steps)
    case(tk.LatticeSpec(size=40, dim=2, defect="wedge"), 4, 8, 0.01)
    for K, steps in [(8, 8), (8, 16), (16, 16), (16, 32)]:
        case(tk.LatticeSpec(size=20, dim=3, layers=5, defect="blackhole", radius=3.3), K, steps)
        case(tk.LatticeSpec(size=20, dim=3, layers=24), K, steps)
    case(tk.LatticeSpec(size=40, dim=3, layers=33, defect="singularity", radius=2.0, mass=2000.0, prune_edges=True), 16, 48, 0.01)
