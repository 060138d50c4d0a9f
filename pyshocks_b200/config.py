"""Process-wide switches of the hot path."""

# arithmetic contract of the kernels launched through the pyshocks-shaped API
# ("fast": re-associated FP64, within 1e-12 of the reference; "strict": the reference's
# operation order, bit-identical to the CPU oracle) -- include/psk.h, enum psk_math
MATH = "fast"


def set_math(mode: str) -> None:
    global MATH
    if mode not in ("fast", "strict"):
        raise ValueError(mode)
    MATH = mode
