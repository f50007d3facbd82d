"""Registry entry `tri_mbt_vsltcls` (class name = args.model.upper(), reference builder/models/__init__.py:15,49)
resolved to the B200-native implementation."""
from medical_tri_modal_pilot_b200.model import TRI_MBT_VSLTCLS  # noqa: F401
