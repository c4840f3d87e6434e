"""shared test helpers (tolerances, synthetic inputs, model/oracle construction)."""
import argparse
import copy

import torch

from dagr_b200.data import EventBatch, format_data, synth_batch
from dagr_b200.utils.args import default_args

# north_star: node features and boxes within 1e-4 relative (fp32).  "relative" is taken against
# |ref| + mean|ref| of the tensor so that exact zeros (ReLU) do not blow up the ratio.
RTOL = 1e-4


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        assert a.numel() == 0
        return 0.0
    s = b.abs().mean().clamp(min=1e-12)
    return float(((a - b).abs() / (b.abs() + s)).max())


def plain_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b|/|b| over the elements with |b| > 1e-3 * mean|b| (plain relative error, no softening term)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0
    m = b.abs() > 1e-3 * b.abs().mean().clamp(min=1e-12)
    if not bool(m.any()):
        return 0.0
    return float(((a - b).abs()[m] / b.abs()[m]).max())


# every assert_close call leaves (softened, plain) max relative errors here; tests/conftest.py dumps the table to
# gpurun_out/parity_errors.json at the end of a GPU session (committed copy: profiles/r02_parity_errors.json)
ERROR_LOG = {}


def assert_close(a, b, tol=RTOL, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    e = rel_err(a, b)
    pe = plain_rel_err(a, b)
    prev = ERROR_LOG.get(what, (0.0, 0.0, 0))
    ERROR_LOG[what] = (max(prev[0], e), max(prev[1], pe), prev[2] + 1)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e} (plain relative, |ref| > 1e-3 mean: {pe:.3e})"


def randomize_bn(model, seed=1):
    """BN running stats / affine randomised so that eval-BN is not the identity."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    # *_pred biases are zero-initialised: make them non-trivial too
    for n, p in model.named_parameters():
        if n.endswith("bias") and "pred" in n:
            with torch.no_grad():
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return model


def make_model(size="s", height=480, width=640, seed=0, dataset="dsec", **over):
    from dagr_b200.model.dagr import DAGR
    torch.manual_seed(seed)
    args = default_args(size, dataset=dataset, **over)
    model = DAGR(args, height=height, width=width).eval()
    randomize_bn(model, seed + 1)
    return model, args


def make_inputs(B, n_events, width, height, seed=42, kind="uniform", ragged=False, window_us=50_000):
    raw = synth_batch(B, n_events, width, height, seed=seed, kind=kind, ragged=ragged, window_us=window_us)
    data = format_data(raw.clone())
    return raw, data
