from sahs_b200.cfgnode import CfgNode  # noqa: F401
