"""Drop-in for the reference's `builder.trainer` package (reference builder/trainer/__init__.py:14-47,
builder/trainer/trainer.py:20-241): same `get_trainer` / `missing_trainer` call contract, B200-native step."""
from medical_tri_modal_pilot_b200.trainer import GradSync, get_trainer, missing_trainer, missing_to_num  # noqa: F401
