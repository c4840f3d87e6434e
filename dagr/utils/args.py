from dagr_b200.utils.args import FLAGS, BASE_FLAGS, parse_config  # noqa: F401
