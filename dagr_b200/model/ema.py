"""Evaluation-side stand-in for the reference's ModelEMA wrapper (src/dagr/model/networks/ema.py:17-33).

scripts/run_test.py:56-62 builds `ModelEMA(model)`, loads `checkpoint['ema']` into `.ema` and evaluates `.ema`;
that is all this class supports: a frozen eval-mode copy of the detector.  The training-time weight averaging of
the reference is out of scope.
"""
import copy


class ModelEMA:
    def __init__(self, model, **_ignored_training_options):
        frozen = copy.deepcopy(model)
        frozen.eval()
        frozen.requires_grad_(False)
        net = getattr(getattr(frozen, "backbone", None), "net", None)
        if net is not None and hasattr(net, "register_hooks"):       # image trunk: hooks do not survive deepcopy
            net.remove_hooks()
            net.register_hooks()
        self.ema = frozen
