from dagr_b200.asynchronous import make_model_asynchronous, AsyncDAGR  # noqa: F401
