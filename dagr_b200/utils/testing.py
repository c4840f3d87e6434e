"""Evaluation loop of the reference's drivers (src/dagr/utils/testing.py), same names and arguments, so that
`scripts/run_test.py` / `run_test_interframe.py`-style code runs unchanged on top of `dagr_b200`:

    to_npy(detections), format_detections(sequences, t, detections)          testing.py:6-14
    run_test_with_visualization(loader, model, dataset, ...)                 testing.py:16-60

wandb image logging (`dagr.utils.logging.log_bboxes`) is used only when that module is importable.  The loop itself is
harness code: all compute happens inside `model(data)`.
"""
from __future__ import annotations

import torch

from .buffers import DetectionBuffer, format_data


def to_npy(detections):
    return [{k: v.cpu().numpy() for k, v in d.items()} for d in detections]


def format_detections(sequences, t, detections):
    out = to_npy(detections)
    for i, det in enumerate(out):
        det["sequence"] = sequences[i]
        det["t"] = t[i]
    return out


def _log_bboxes():
    try:
        from dagr.utils.logging import log_bboxes             # the reference's wandb logger, if present
        return log_bboxes
    except Exception:
        return None


def run_test_with_visualization(loader, model, dataset: str, log_every_n_batch=-1, name="", compile_detections=False,
                                no_eval=False):
    model.eval()
    mapcalc = None
    if not no_eval:
        mapcalc = DetectionBuffer(height=loader.dataset.height, width=loader.dataset.width, classes=loader.dataset.classes)
    compiled = [] if compile_detections else None
    log_bboxes = _log_bboxes() if log_every_n_batch > 0 else None
    on_gpu = torch.cuda.is_available()
    for i, data in enumerate(loader):
        if on_gpu:
            data = data.cuda(non_blocking=True)
        shown = data.clone() if log_bboxes is not None else None
        data = format_data(data)
        detections, targets = model(data.clone())
        if compile_detections:
            compiled.extend(format_detections(data.sequence, data.t1, detections))
        if log_bboxes is not None and i % log_every_n_batch == 0:
            log_bboxes(shown, targets=targets, detections=detections, bidx=4, class_names=loader.dataset.classes,
                       key="testing/evaluated_bboxes")
        if mapcalc is not None:
            mapcalc.update(detections, targets, dataset, data.height[0], data.width[0])
    result = mapcalc.compute() if mapcalc is not None else None
    return (result, compiled) if compile_detections else result
