from sahs_b200._loaders import load_llff_data  # noqa: F401
