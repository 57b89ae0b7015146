from sahs_b200._loaders import load_blender_data  # noqa: F401
