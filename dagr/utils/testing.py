from dagr_b200.utils.testing import format_detections, run_test_with_visualization, to_npy  # noqa: F401
