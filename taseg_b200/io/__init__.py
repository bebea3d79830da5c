"""On-disk formats that feed the hot path: SemanticKITTI sequences, nuScenes sweep metadata."""
from .kitti import (KittiSequence, load_sample, parse_calibration, parse_poses, read_labels, read_scan,  # noqa: F401
                    write_labels)
from .nuscenes import quaternion_rotation_matrix, read_sweep, select_sweeps, sweep_dt, sweep_transform  # noqa: F401,E402
