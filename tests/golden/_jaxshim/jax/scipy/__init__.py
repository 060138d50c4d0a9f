"""``jax.scipy`` facade (fixture tooling only)."""
from . import special  # noqa: F401
