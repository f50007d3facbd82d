"""ORACLE SUPPORT (test infrastructure): the synthetic-batch generator lives in the package
(medical_tri_modal_pilot_b200/synth.py, it is also bench.py's workload); re-exported here for the tests."""
from medical_tri_modal_pilot_b200.synth import make_batch  # noqa: F401
