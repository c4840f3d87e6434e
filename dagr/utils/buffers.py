from dagr_b200.utils.buffers import (Buffer, DetectionBuffer, DictBuffer, bbox_t_to_ndarray, compile, diag_filter,  # noqa: F401
                                     filter_bboxes, format_data, records_from_device, save_detections, to_cpu, to_npy)
