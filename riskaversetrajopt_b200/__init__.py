"""B200-native SAA linearize+assemble path for risk-averse trajectory optimization.

Host side mirrors the ``Model`` classes of StanfordASL/RiskAverseTrajOpt
(drone/drone_risk.py, car/driving.py, hopper/hopper.py); the sample-parallel
work runs in hand-written sm_100a CUDA behind the C-ABI of ``libsaa_b200.so``
(see ``include/saa_b200.h``).  There is no CPU fallback: constructing a model
without the built library or without a CUDA device raises.
"""
__version__ = "0.1.0"
