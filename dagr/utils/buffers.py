from dagr_b200.data import format_data  # noqa: F401
