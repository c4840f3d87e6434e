from dagr_b200.model.net import Net  # noqa: F401
