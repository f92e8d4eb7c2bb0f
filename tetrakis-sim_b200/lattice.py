"""Lattice specification: a tetrakis sheet + defect described by parameters instead of a graph.

The reference builds every lattice as a ``networkx`` graph (``tetrakis_sim/lattice.py:15-76``),
which is O(N * size) edges -- 2.2e12 edges for BASELINE config 2 -- so the large configurations
cannot exist as graphs at all.  A :class:`LatticeSpec` carries the same
information the reference's builders take as arguments and produces, in O(defect volume), the
sparse description the device plan needs: which nodes a defect removed / pruned / tagged and
which single edges it cut.  ``run_wave_sim`` accepts a ``LatticeSpec`` wherever it accepts a
graph.

Semantics restated from the reference (file:line relative to /root/reference/tetrakis_sim):
  * node tuples ``(r, c, q)`` / ``(r, c, z, q)``, ``q`` in "ABCD"      lattice.py:30, 48-54
  * black hole: remove nodes with dist <  radius                       defects.py:110-135
      (2-tuple centre: ``math.hypot`` in (r, c), a cylinder through every layer; 3-tuple: sqrt)
  * singularity: tag nodes with dist <= radius with mass/potential,    defects.py:194-247
      ``prune_edges`` strips every edge incident to a tagged node
  * wedge: cut the single edge A-B of the centre cell (on one layer)   defects.py:55-107
  * default centre (size//2, size//2[, layers//2])                     batch.py:89-92
  * kick rule                                                          batch.py:124-151
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

QUADRANTS = "ABCD"
FLAG_ALIVE, FLAG_PART = 1, 2


@dataclass
class LatticeSpec:
    size: int
    dim: int = 2
    layers: int = 1
    defect: str = "none"  # none | blackhole | singularity | wedge
    center: tuple | None = None  # defect centre; default = lattice centre (batch.py:89-92)
    radius: float = 2.5
    mass: float = 1000.0
    potential: float = 0.0
    prune_edges: bool = False
    wedge_layer: int | None = None
    _cache: dict = field(default_factory=dict, repr=False, compare=False)

    def __post_init__(self):
        if self.dim not in (2, 3):
            raise ValueError("dim must be 2 or 3")
        if self.dim == 2:
            self.layers = 1
        if self.size < 1 or self.layers < 1:
            raise ValueError("size and layers must be >= 1")
        if self.defect not in ("none", "blackhole", "singularity", "wedge"):
            raise ValueError(f"Unsupported defect type '{self.defect}'")

    # ------------------------------------------------------------------ geometry
    @property
    def plane(self) -> int:
        return self.size * self.size * 4

    @property
    def num_slots(self) -> int:
        """Node slots of the full lattice (removed nodes included)."""
        return self.layers * self.plane

    @property
    def lattice_center(self) -> tuple:
        s = self.size // 2
        return (s, s) if self.dim == 2 else (s, s, self.layers // 2)

    @property
    def defect_center(self) -> tuple:
        return tuple(self.center) if self.center is not None else self.lattice_center

    def index(self, node) -> int:
        """Reference node tuple -> global linear index ((z*size + r)*size + c)*4 + q."""
        if len(node) == 3:
            r, c, q = node
            z = 0
        else:
            r, c, z, q = node
        qi = QUADRANTS.index(q) if isinstance(q, str) else int(q)
        if not (0 <= r < self.size and 0 <= c < self.size and 0 <= z < self.layers):
            raise KeyError(node)
        return ((z * self.size + r) * self.size + c) * 4 + qi

    def node(self, index: int) -> tuple:
        q = index & 3
        cell = index >> 2
        c = cell % self.size
        r = (cell // self.size) % self.size
        z = cell // (self.size * self.size)
        return (r, c, QUADRANTS[q]) if self.dim == 2 else (r, c, z, QUADRANTS[q])

    # ------------------------------------------------------------------ defects
    def _cells_within(self, inclusive: bool):
        """Cells (z, r, c) whose distance to the defect centre is < radius (or <= radius), using
        the reference's own float expressions on integer offsets (defects.py:120-131, 214-219)."""
        key = ("cells", inclusive)
        if key in self._cache:
            return self._cache[key]
        ctr = self.defect_center
        R = float(self.radius)
        if math.isnan(R) or R < 0 or (R == 0 and not inclusive):
            out = np.empty((0, 3), dtype=np.int64)
            self._cache[key] = out
            return out
        span = int(math.floor(R)) + 1
        r0, c0 = ctr[0], ctr[1]
        rs = [r for r in range(int(math.floor(r0 - span)), int(math.ceil(r0 + span)) + 1) if 0 <= r < self.size]
        cs = [c for c in range(int(math.floor(c0 - span)), int(math.ceil(c0 + span)) + 1) if 0 <= c < self.size]
        hits = []
        if len(ctr) == 2:
            ok2 = []
            for r in rs:
                for c in cs:
                    d = math.hypot(r - r0, c - c0)
                    if d < R or (inclusive and d <= R):
                        ok2.append((r, c))
            for z in range(self.layers):
                hits.extend((z, r, c) for r, c in ok2)
        else:
            z0 = ctr[2]
            zs = [z for z in range(int(math.floor(z0 - span)), int(math.ceil(z0 + span)) + 1) if 0 <= z < self.layers]
            if len(rs) * len(cs) * len(zs) > 200_000:
                # vectorised: sqrt of an exactly representable sum is correctly rounded in both
                # numpy and math.sqrt, so the comparison is identical
                zz, rr, cc = np.meshgrid(np.array(zs), np.array(rs), np.array(cs), indexing="ij")
                d = np.sqrt(((rr - r0) ** 2 + (cc - c0) ** 2 + (zz - z0) ** 2).astype(np.float64))
                m = (d <= R) if inclusive else (d < R)
                out = np.stack([zz[m], rr[m], cc[m]], axis=1).astype(np.int64)
                self._cache[key] = out
                return out
            for z in zs:
                for r in rs:
                    for c in cs:
                        d = math.sqrt((r - r0) ** 2 + (c - c0) ** 2 + (z - z0) ** 2)
                        if d < R or (inclusive and d <= R):
                            hits.append((z, r, c))
        out = np.array(hits, dtype=np.int64).reshape(-1, 3)
        self._cache[key] = out
        return out

    def specials(self):
        """(index[int64], flags[uint8], inv_mass[f64], potential[f64]) of every node the defect
        touched -- the ``special_*`` arrays of ``tk_lattice_desc``."""
        if "specials" in self._cache:
            return self._cache["specials"]
        S = self.size
        if self.defect == "blackhole":
            cells = self._cells_within(inclusive=False)
            flags, im, pot = 0, 1.0, 0.0
        elif self.defect == "singularity":
            cells = self._cells_within(inclusive=True)
            flags = FLAG_ALIVE | (0 if self.prune_edges else FLAG_PART)
            m = float(self.mass)
            im = 1.0 / (m if m > 0.0 else 1.0)  # physics.py:98-101
            pot = float(self.potential)
        else:
            cells = np.empty((0, 3), dtype=np.int64)
            flags, im, pot = FLAG_ALIVE | FLAG_PART, 1.0, 0.0
        base = ((cells[:, 0] * S + cells[:, 1]) * S + cells[:, 2]) * 4
        idx = (base[:, None] + np.arange(4)[None, :]).reshape(-1)
        out = (
            idx.astype(np.int64),
            np.full(idx.size, flags, dtype=np.uint8),
            np.full(idx.size, im, dtype=np.float64),
            np.full(idx.size, pot, dtype=np.float64),
        )
        self._cache["specials"] = out
        return out

    def corrections(self):
        """Directed edge corrections (a, b, sign) relative to the clique model: the wedge cut."""
        if self.defect != "wedge":
            return []
        ctr = self.defect_center
        r, c = int(ctr[0]), int(ctr[1])
        if not (0 <= r < self.size and 0 <= c < self.size):
            return []
        if self.dim == 2:
            z = 0
        else:
            z = self.wedge_layer if self.wedge_layer is not None else self.lattice_center[2]
            if not (0 <= z < self.layers):
                raise ValueError(f"Layer {z} is outside of lattice range")
        a = ((z * self.size + r) * self.size + c) * 4
        return [(a, a + 1, -1.0), (a + 1, a, -1.0)]

    # ------------------------------------------------------------------ bookkeeping
    @property
    def removed_count(self) -> int:
        return int(self.specials()[0].size) if self.defect == "blackhole" else 0

    @property
    def num_nodes(self) -> int:
        """Nodes that exist in the graph (what the reference's G.number_of_nodes() returns)."""
        return self.num_slots - self.removed_count

    def _blocked(self):
        """Sets of removed / unkickable node indices (small: defect volume)."""
        if "blocked" not in self._cache:
            idx, flags, _, _ = self.specials()
            removed = set(idx[(flags & FLAG_ALIVE) == 0].tolist())
            unkick = set(idx.tolist()) if self.defect == "singularity" else set(removed)
            self._cache["blocked"] = (removed, unkick)
        return self._cache["blocked"]

    def contains(self, node) -> bool:
        try:
            i = self.index(node)
        except (KeyError, ValueError, TypeError):
            return False
        return i not in self._blocked()[0]

    def default_node_index(self) -> int:
        """Index of ``nodes[len(nodes)//2]`` in the reference's node order (physics.py:37) --
        r-major, then c, z, q (lattice.py:30,48-54) over the nodes that exist."""
        if self.removed_count:
            removed = self._blocked()[0]
            k = self.num_nodes // 2
            # walk the reference order, skipping removed nodes (rare path, O(N) worst case)
            seen = 0
            for r in range(self.size):
                for c in range(self.size):
                    for z in range(self.layers):
                        for q in range(4):
                            i = ((z * self.size + r) * self.size + c) * 4 + q
                            if i in removed:
                                continue
                            if seen == k:
                                return i
                            seen += 1
            raise RuntimeError("empty lattice")
        k = self.num_slots // 2
        q = k % 4
        k //= 4
        z = k % self.layers
        k //= self.layers
        c = k % self.size
        r = k // self.size
        return ((z * self.size + r) * self.size + c) * 4 + q

    def kick_node(self) -> tuple:
        """batch.py:124-151: the connected, non-singular surviving node of the central layer that is
        closest (in r, c) to the lattice centre; ties go to the first in node order (r, c, q)."""
        ctr = self.lattice_center
        z = ctr[2] if self.dim == 3 else 0
        removed, unkick = self._blocked()
        for blocked in (unkick, removed):  # fall back to any surviving node (batch.py:150)
            best = None
            reach = int(math.ceil(self.radius)) + 2 if self.defect in ("blackhole", "singularity") else 1
            while best is None and reach <= 2 * self.size + 2:
                for r in range(max(0, ctr[0] - reach), min(self.size, ctr[0] + reach + 1)):
                    for c in range(max(0, ctr[1] - reach), min(self.size, ctr[1] + reach + 1)):
                        d2 = (r - ctr[0]) ** 2 + (c - ctr[1]) ** 2
                        if d2 > reach * reach:
                            continue  # only trust candidates inside the searched disc
                        for q in range(4):
                            i = ((z * self.size + r) * self.size + c) * 4 + q
                            if i in blocked:
                                continue
                            key = (d2, r, c, q)
                            if best is None or key < best:
                                best = key
                reach *= 2
            if best is not None:
                _, r, c, q = best
                return (r, c, QUADRANTS[q]) if self.dim == 2 else (r, c, z, QUADRANTS[q])
        raise RuntimeError("No nodes remain on the central layer after defect application")

    def bench_kick_node(self) -> tuple:
        """scripts/bench_wave.py:43-46."""
        s = self.size // 2
        return (s, s, "A") if self.dim == 2 else (s, s, self.layers // 2, "A")

    def to_networkx(self):
        """The lattice + defect as a ``networkx`` graph, for sizes networkx can hold (plotting,
        curvature and the other graph consumers of the reference keep working).  Node order, adjacency
        order and attributes are those of ``apply_defect(build_sheet(...))``: edges are inserted layer
        by layer as cell cliques, row cliques, column cliques, then the vertical links
        (lattice.py:37-44, 58-72) -- skipping what the defect removed, which leaves the same relative
        order as removing it afterwards."""
        import networkx as nx
        from itertools import combinations

        S, L = self.size, self.layers
        idx, flags, im, pot = self.specials()
        dead = set(idx[(flags & FLAG_ALIVE) == 0].tolist())
        loose = set(idx[(flags & FLAG_PART) == 0].tolist()) | dead  # nodes without any edge
        cut = {(a, b) for a, b, _ in self.corrections()}

        def nid(z, r, c, q):
            return ((z * S + r) * S + c) * 4 + q

        def label(i):
            return self.node(i)

        G = nx.Graph()
        order = (nid(z, r, c, q) for r in range(S) for c in range(S) for z in range(L) for q in range(4))
        G.add_nodes_from(label(i) for i in order if i not in dead)

        def link(a, b):
            if a not in loose and b not in loose and (a, b) not in cut:
                G.add_edge(label(a), label(b))

        for z in range(L):
            for r in range(S):
                for c in range(S):
                    for qa, qb in combinations(range(4), 2):
                        link(nid(z, r, c, qa), nid(z, r, c, qb))
            for r in range(S):
                for q in range(4):
                    for ca, cb in combinations(range(S), 2):
                        link(nid(z, r, ca, q), nid(z, r, cb, q))
            for c in range(S):
                for q in range(4):
                    for ra, rb in combinations(range(S), 2):
                        link(nid(z, ra, c, q), nid(z, rb, c, q))
        for z in range(L - 1):
            for r in range(S):
                for c in range(S):
                    for q in range(4):
                        link(nid(z, r, c, q), nid(z + 1, r, c, q))
        if self.defect == "singularity":
            for i in idx.tolist():
                if i not in dead:
                    G.nodes[label(i)].update(mass=float(self.mass), potential=float(self.potential),
                                             singular=True)
        return G

    def slab(self, rank: int, world: int) -> tuple[int, int]:
        """Contiguous block of layers owned by ``rank`` of ``world``."""
        base, extra = divmod(self.layers, world)
        lo = rank * base + min(rank, extra)
        return lo, lo + base + (1 if rank < extra else 0)
