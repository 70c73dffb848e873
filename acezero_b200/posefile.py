"""ACE pose-file reader / writer, byte-compatible with the reference's `dataset_io` (reference dataset_io.py:96-186):
one line per image `file qw qx qy qz tx ty tz focal confidence`, world-to-camera, OpenCV convention."""
import numpy as np
import torch


def _quat_xyzw_from_matrix(R):
    """scipy.spatial.transform.Rotation.from_matrix(R).as_quat() (x, y, z, w); scipy is used when available so that
    the printed digits match the reference exactly."""
    try:
        from scipy.spatial.transform import Rotation
        return Rotation.from_matrix(np.asarray(R, dtype=np.float64)).as_quat()
    except ImportError:  # pragma: no cover
        R = np.asarray(R, dtype=np.float64)
        t = np.trace(R)
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
        else:
            i = int(np.argmax(np.diag(R)))
            j, k = (i + 1) % 3, (i + 2) % 3
            s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
            q = [0, 0, 0, (R[k, j] - R[j, k]) / s]
            q[i], q[j], q[k] = 0.25 * s, (R[j, i] + R[i, j]) / s, (R[k, i] + R[i, k]) / s
        return np.array(q)


def write_pose_to_pose_file(out_pose_file, rgb_file, pose, confidence, focal_length):
    """reference dataset_io.py:159-186. pose: numpy 4x4 / 3x4 world-to-cam."""
    q = _quat_xyzw_from_matrix(pose[:3, :3])
    t = pose[:3, 3]
    out_pose_file.write(f"{rgb_file} {q[3]} {q[0]} {q[1]} {q[2]} {t[0]} {t[1]} {t[2]} {focal_length} {confidence}\n")


def load_dataset_ace(pose_file, confidence_threshold):
    """reference dataset_io.py:96-156: returns (rgb_files, cam-to-world 4x4 float tensors, focal lengths)."""
    from scipy.spatial.transform import Rotation
    rgb_files, poses, focal_lengths = [], [], []
    with open(pose_file, 'r') as f:
        for line in f.readlines():
            tok = line.split()
            assert len(tok) == 10, f"Expected 10 tokens per line in pose file, got {len(tok)}"
            if float(tok[-1]) < confidence_threshold:
                continue
            q_wxyz = [float(t) for t in tok[1:5]]
            T = np.eye(4)
            T[:3, :3] = Rotation.from_quat(q_wxyz[1:] + [q_wxyz[0]]).as_matrix()
            T[:3, 3] = [float(t) for t in tok[5:8]]
            rgb_files.append(tok[0])
            focal_lengths.append(float(tok[-2]))
            poses.append(torch.from_numpy(np.linalg.inv(T)).float())
    return rgb_files, poses, focal_lengths
