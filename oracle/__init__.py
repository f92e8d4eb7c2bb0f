"""CPU oracle for the tetrakis-sim wave-propagation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the CPU arm being timed), never on the CUDA product path.

What is restated here (all citations relative to the reference checkout):

* ``lattice_oracle``  - ``tetrakis_sim/lattice.py:15-76`` (graph construction,
  including node and adjacency insertion order) and ``tetrakis_sim/defects.py``
  ``:55-107`` (wedge), ``:110-135`` (black hole), ``:194-247`` (singularity),
  plus the kick rule of ``tetrakis_sim/batch.py:124-151``.
* ``wave_oracle``     - ``tetrakis_sim/physics.py:41-131`` (CFL clamp, attribute
  precompute, leapfrog loop in the reference's exact operation order) and
  ``physics.py:140-165`` (probe FFT), in pure Python, NumPy and plain C
  (``wave_oracle.c``); and the rank-structured "clique-sum" restatement of the
  same operator used for lattices far too large for networkx.

Parity pinning: the reference ships no numeric golden vectors for u(t)
(SURVEY.md section 0.4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` and stores graphs, full histories, spectra and metadata in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` requires the generic
restatement to reproduce those histories BIT-EXACTLY and the clique-sum
restatement to within 1e-13 of max|u|.
"""
