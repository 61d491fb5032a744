"""CPU oracles for the SAA linearize+assemble path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.
The product path (``riskaversetrajopt_b200``) never falls back to these
functions; it raises if ``libsaa_b200.so`` is missing.

PARITY PINNING.  The reference (StanfordASL/RiskAverseTrajOpt) ships no tests,
golden vectors or stored results, and cannot be imported as is in this image (it
needs ``jax``/``jaxlib``/``osqp``/``ipyopt``/``matplotlib``, none installed, no
network).  It is pure Python, though, and the only thing the model code needs from
JAX is mechanism (array ops + autodiff).  ``oracle/refexec`` supplies that mechanism
(``minijax``: NumPy + exact forward-mode dual numbers, unit-tested against closed
forms, finite differences and mpmath) and EXECUTES THE REFERENCE'S OWN SOURCE, read
from /root/reference where it lies (nothing is copied).  What pins this oracle:

* ``tests/golden/ref_*.npz`` are outputs of the reference itself run here:
  ``tests/golden/make_golden_ref.py`` (committed) calls the reference's
  ``Model.get_constraints_coeffs`` / ``us_to_state_trajectories`` / Monte-Carlo
  verification functions / ``slip_risk_constraints`` + ``jacrev`` + ``hessian`` at
  the reference's seeds (drone/drone_risk.py:57, car/driving.py:61,
  hopper/hopper.py:33), for both drone scripts (drone_risk.py, drone_times.py),
  both methods, relaxed and normal iterations, and the car's relaxation edge cases.
  ``tests/test_reference_exec.py`` re-executes the reference in this container and
  requires the committed fixtures to be reproduced exactly.
* ``oracle_a`` restates the reference line by line with a second, independent
  autodiff (``torch.func.vmap(jacfwd(...))``, float64); ``tests/golden/*.npz``
  (without the ``ref_`` prefix) are minted from it.  It agrees with the reference's
  own execution to <= 4e-13 relative, patterns bit-exact.
* ``oracle_b`` is an analytic closed-form restatement that scales to 10^6+
  samples, ``saa_oracle.c`` its C/OpenMP port (CPU baseline); both are tested
  against the reference-minted fixtures.
"""
