"""Drop-in for the reference's masking_generator.py: same names, B200 implementation."""
from mofo_b200.masking_generator import TubeMaskingGenerator, TubeMaskingGenerator_BB  # noqa: F401
