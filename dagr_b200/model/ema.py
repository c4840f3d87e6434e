"""ModelEMA (reference: src/dagr/model/networks/ema.py:6-51), eval-side surface only."""
import math
from copy import deepcopy

import torch


class ModelEMA:
    def __init__(self, model, decay=0.9999, updates=0):
        self.ema = deepcopy(model).eval()
        try:
            self.ema.backbone.net.remove_hooks()
            self.ema.backbone.net.register_hooks()
        except Exception:
            pass
        self.updates = updates
        self.decay = lambda x: decay * (1 - math.exp(-x / 2000))
        for p in self.ema.parameters():
            p.requires_grad_(False)

    def update(self, model):
        with torch.no_grad():
            self.updates += 1
            d = self.decay(self.updates)
            msd = model.state_dict()
            for k, v in self.ema.state_dict().items():
                if v.dtype.is_floating_point:
                    v *= d
                    v += (1.0 - d) * msd[k].detach()
