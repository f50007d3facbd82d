"""Drop-in shim with the reference's package path (`builder.models.get_model`, reference builder/models/__init__.py)."""
