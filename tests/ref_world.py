"""Test helper: run the UNMODIFIED reference classes on a synthetic world (the code lives in oracle/ref_harness.py so that
bench.py's baseline legs can use it as well)."""
from ref_harness import make_fake_task, reference_assets, reference_step  # noqa: F401
