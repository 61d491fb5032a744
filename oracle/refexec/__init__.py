"""Reference executor: runs the reference's own source (read from /root/reference) on a NumPy
stand-in for JAX.  TEST INFRASTRUCTURE ONLY -- see loader.py / minijax.py."""
from .loader import load_script, load_nested, extract_functions, reference_available, REFERENCE_ROOT  # noqa: F401
