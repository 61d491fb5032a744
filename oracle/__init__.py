"""CPU oracles for the SAA linearize+assemble path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.
The product path (``riskaversetrajopt_b200``) never falls back to these
functions; it raises if ``libsaa_b200.so`` is missing.

PARITY PINNING.  The reference (StanfordASL/RiskAverseTrajOpt) ships no tests,
golden vectors or stored results, and cannot be imported in this image (needs
``jax``/``jaxlib``/``osqp``/``ipyopt``, none installed, no network; it also uses
APIs removed from current JAX/NumPy).  **Parity is therefore unpinned by the
reference itself.**  What pins this oracle instead:

* ``oracle_a`` restates the reference line by line with an *independent*
  autodiff (``torch.func.vmap(jacfwd(...))``, float64) – the same mechanism the
  reference uses (``jax.vmap(jax.jacfwd(...))``) applied to the same in-tree
  formulas, including the dense packing and the SciPy CSR/CSC tail.
* ``oracle_b`` is an analytic closed-form restatement that scales to 10^6+
  samples; it must agree with ``oracle_a`` to <= 1e-12 (tests/test_oracle.py)
  and with central finite differences of the rollout.
* inputs are pinned by the reference's seeds and RNG call order
  (drone/drone_risk.py:57, car/driving.py:61, hopper/hopper.py:33).
* ``tests/golden/*.npz`` are minted from ``oracle_a`` by
  ``tests/golden/make_golden.py`` (committed).
"""
