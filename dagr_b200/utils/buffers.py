"""Host-side detection records: what the reference does with the model's output on the CPU.

Mirrors, with the same names and argument meaning:
  src/dagr/utils/buffers.py:10-30    diag_filter / filter_bboxes
  src/dagr/utils/buffers.py:46-76    bbox_t_to_ndarray / compile  (record dtype t,x,y,w,h,class_id[,class_confidence])
  src/dagr/utils/buffers.py:78-124   to_cpu / Buffer / DetectionBuffer (compile, update; compute needs the COCO tools)
  src/dagr/utils/buffers.py:127-145  DictBuffer
  scripts/run_test_interframe.py:21-45  to_npy / save_detections
and adds `records_from_device`, which turns the batched device output of `Engine.postprocess` (det [B,A,6], ndet [B])
into one record array per image with a single device->host copy (the reference loops over images and tensors).
COCO evaluation itself (pycocotools, src/dagr/utils/coco_eval.py) is outside the hot path and not rebuilt.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Sequence

import numpy as np
import torch

from ..data import format_data  # noqa: F401  (re-exported like the reference module)

_REC = [("t", "<u8"), ("x", "<f4"), ("y", "<f4"), ("w", "<f4"), ("h", "<f4"), ("class_id", "u1")]
_REC_CONF = _REC + [("class_confidence", "<f4")]


def diag_filter(bbox: torch.Tensor, height: int, width: int, min_box_diagonal: int = 30, min_box_side: int = 20):
    """clamps the xyxy boxes to the image IN PLACE (like the reference) and returns the keep mask."""
    bbox[..., 0::2] = torch.clamp(bbox[..., 0::2], 0, width - 1)
    bbox[..., 1::2] = torch.clamp(bbox[..., 1::2], 0, height - 1)
    wh = bbox[..., 2:] - bbox[..., :2]
    w, h = wh[..., 0], wh[..., 1]
    diag = torch.sqrt(w ** 2 + h ** 2)
    return (diag > min_box_diagonal) & (w > min_box_side) & (h > min_box_side)


def filter_bboxes(detections: List[Dict[str, torch.Tensor]], height: int, width: int, min_box_diagonal: int = 30,
                  min_box_side: int = 20):
    out = []
    for d in detections:
        mask = diag_filter(d["boxes"], height, width, min_box_diagonal, min_box_side)
        out.append({k: v[mask] for k, v in d.items()})
    return out


def _records(t, boxes: np.ndarray, labels: np.ndarray, scores=None) -> np.ndarray:
    rec = np.zeros(shape=(len(boxes),), dtype=_REC if scores is None else _REC_CONF)
    rec["t"] = t
    rec["x"], rec["y"] = boxes[:, 0], boxes[:, 1]
    rec["w"], rec["h"] = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    rec["class_id"] = labels
    if scores is not None:
        rec["class_confidence"] = scores
    return rec


def bbox_t_to_ndarray(bbox: Dict[str, torch.Tensor], t) -> np.ndarray:
    """one image: dict(boxes xyxy, labels[, scores]) on the CPU -> structured records; a dict with three entries is a
    detection (gets class_confidence), one with two a ground-truth box set."""
    scores = bbox["scores"].numpy() if len(bbox) == 3 else None
    return _records(t, bbox["boxes"].numpy(), bbox["labels"].numpy(), scores)


def compile(detections, sequences, timestamps):  # noqa: A001  (name kept from the reference)
    per_seq: Dict[str, list] = {}
    for det, s, t in zip(detections, sequences, timestamps):
        per_seq.setdefault(s, []).append(bbox_t_to_ndarray(det, t))
    return {k: np.concatenate(v) for k, v in per_seq.items() if len(v) > 0}


def to_cpu(data_list: List[Dict[str, torch.Tensor]]):
    return [{k: v.cpu() for k, v in d.items()} for d in data_list]


class Buffer:
    def __init__(self):
        self.buffer = []

    def extend(self, elements: List[Dict[str, torch.Tensor]]):
        self.buffer.extend(to_cpu(elements))

    def clear(self):
        self.buffer.clear()

    def __iter__(self):
        return iter(self.buffer)

    def __len__(self):
        return len(self.buffer)


class DetectionBuffer:
    def __init__(self, height: int, width: int, classes: List[str]):
        self.height, self.width, self.classes = height, width, classes
        self.detections, self.ground_truth = Buffer(), Buffer()

    def compile(self, sequences, timestamps):
        return compile(self.detections, sequences, timestamps), compile(self.ground_truth, sequences, timestamps)

    def update(self, detections, groundtruth, dataset: str = "", height=None, width=None):
        self.detections.extend(detections)
        self.ground_truth.extend(groundtruth)

    def compute(self) -> Dict[str, float]:
        try:
            from dagr.utils.coco_eval import evaluate_detection              # the reference's own evaluator, if installed
        except Exception as e:                                               # pragma: no cover
            raise NotImplementedError("mAP needs the reference's COCO evaluation (pycocotools); this build stops at the "
                                      "detection records (compile / save_detections)") from e
        out = evaluate_detection(self.ground_truth.buffer, self.detections.buffer, height=self.height, width=self.width,
                                 classes=self.classes)
        out = {k.replace("AP", "mAP"): v for k, v in out.items()}
        self.detections.clear()
        self.ground_truth.clear()
        return out


class DictBuffer:
    """running mean of dictionaries of floats (scripts/count_flops.py)."""

    def __init__(self):
        self.running_mean, self.n = None, 0

    def update(self, dictionary: Dict[str, float]):
        if self.running_mean is None:
            self.running_mean = {k: 0 for k in dictionary}
        n = self.n
        self.running_mean = {k: n / (n + 1) * self.running_mean[k] + dictionary[k] / (n + 1) for k in dictionary}
        self.n += 1

    def save(self, path):
        torch.save(self.running_mean, path)

    def compute(self) -> Dict[str, float]:
        return self.running_mean


# ---- scripts/run_test_interframe.py:21-45 ---------------------------------------------------------------------------
def to_npy(detections) -> np.ndarray:
    """dict(boxes, labels, scores, t) of one image -> records with class_confidence (w, h derived from the stored x, y)."""
    boxes = np.asarray(detections["boxes"], dtype=np.float32).reshape(-1, 4)
    rec = np.zeros(shape=(len(boxes),), dtype=np.dtype(_REC_CONF))
    rec["t"] = detections["t"]
    rec["x"], rec["y"] = boxes[:, 0], boxes[:, 1]
    rec["w"], rec["h"] = boxes[:, 2] - rec["x"], boxes[:, 3] - rec["y"]
    rec["class_id"] = np.asarray(detections["labels"])
    rec["class_confidence"] = np.asarray(detections["scores"])
    return rec


def save_detections(directory, detections: Sequence[dict]):
    """groups by d['sequence'], sorts each sequence by t and writes detections_<sequence>.npy; returns the arrays."""
    per_seq: Dict[str, np.ndarray] = {}
    for d in detections:
        s, rec = d["sequence"], to_npy(d)
        per_seq[s] = rec if s not in per_seq else np.concatenate([per_seq[s], rec])
    out = {}
    for s, rec in per_seq.items():
        rec = rec[rec["t"].argsort()]
        np.save(Path(directory) / f"detections_{s}.npy", rec)
        out[s] = rec
    return out


# ---- batched device output -> records ------------------------------------------------------------------------------
def records_from_device(det: torch.Tensor, ndet: torch.Tensor, timestamps: Sequence[int]) -> List[np.ndarray]:
    """det float32[B,A,6] = (x1,y1,x2,y2,score,label) sorted by score, ndet int32[B] (Engine.postprocess / the batched
    NMS kernel) -> one record array per image, with ONE device->host copy for the whole batch."""
    B, A, _ = det.shape
    host = torch.cat([det.reshape(B, -1), ndet.reshape(B, 1).to(det.dtype)], dim=1).cpu().numpy()
    out = []
    for b in range(B):
        n = int(host[b, -1])
        d = host[b, : A * 6].reshape(A, 6)[:n]
        out.append(_records(timestamps[b], d[:, :4], d[:, 5].astype(np.uint8), d[:, 4]))
    return out
