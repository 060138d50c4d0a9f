"""``jax.scipy.special`` facade (fixture tooling only)."""
from scipy.special import erf  # noqa: F401
