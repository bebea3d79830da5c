"""On-disk formats that feed the hot path: SemanticKITTI sequences."""
from .kitti import (KittiSequence, load_sample, parse_calibration, parse_poses, read_labels, read_scan,  # noqa: F401
                    write_labels)
