"""b200slam: B200-native ALIKED + LightGlue frontend (see ../b200slam/__init__.py for the import shim)."""
