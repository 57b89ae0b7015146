from sahs_b200._loaders import load_flame_data  # noqa: F401
