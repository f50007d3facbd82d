"""Model registry with the reference's contract (reference builder/models/__init__.py:14-51):
`get_model(args)` imports `builder.models.8_missing_models.<args.model>` and returns the class `<args.model>.upper()`.
Only `tri_mbt_vsltcls` is provided by the B200 build; any other name raises like the reference does for unknown models."""
import importlib


def get_model(args):
    model_module = importlib.import_module("builder.models.8_missing_models." + args.model)
    return getattr(model_module, args.model.upper())
