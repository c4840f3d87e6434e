from dagr_b200.model.dagr import DAGR, GNNHead, CNNHead  # noqa: F401
