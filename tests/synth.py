"""Re-export of the product's synthetic-workload helper for the tests."""
from apm_b200.synth import TEXT_SEED, make_patterns, text_slice  # noqa: F401
