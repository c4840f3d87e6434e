from dagr_b200.model.ema import ModelEMA  # noqa: F401
