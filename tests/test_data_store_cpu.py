"""Host logic of the offline painter's store (lav_b200/data_paint.py): the reference's key layout
(lav/utils/datasets/basic_dataset.py:52-53,82-101; point_paint_dataset.py:34-46) round-trips through the directory environment."""
import os

import numpy as np
import pytest


def test_dir_env_keeps_the_reference_key_layout(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from lav_b200.data_paint import DirEnv, PointPaintDataset
    rs = np.random.RandomState(0)
    frames = []
    for t, n_frames in enumerate((2, 3)):
        env = DirEnv(str(tmp_path / f"t{t}"))
        env.put("len", str(n_frames).encode())
        env.put("town", b"Town03")
        for i in range(n_frames):
            lidar = rs.randn(100 + 7 * i, 4).astype(np.float32)
            rgb = rs.randint(0, 256, (3, 288, 256, 3), dtype=np.uint8)          # BGR as cv2 stores it
            env.put(f"lidar_{i:05d}", lidar.tobytes())
            for c in range(3):
                env.put(f"rgb_{c}_{i:05d}", cv2.imencode(".png", rgb[c])[1].tobytes())
            frames.append((lidar, rgb))
    (tmp_path / "not_a_trajectory.txt").write_text("x")
    ds = PointPaintDataset(str(tmp_path))
    assert len(ds) == 5
    for idx, (lidar, rgb) in enumerate(frames):
        got_l, got_rgb = ds[idx]
        assert np.array_equal(got_l, lidar)
        assert got_rgb.shape == (3, 3, 288, 256) and np.array_equal(got_rgb, rgb[..., ::-1].transpose(0, 3, 1, 2))     # RGB, NCHW
    painted = rs.rand(len(frames[3][0]), 4)
    ds.commit(3, painted)
    raw = DirEnv(str(tmp_path / "t1")).get("lidar_sem_00001")               # frame 3 = trajectory 1, index 1
    assert np.array_equal(np.frombuffer(raw, np.float32).reshape(-1, 4), painted.astype(np.float32))
    assert sorted(os.listdir(tmp_path / "t1" / "kv"))[:3] == ["len", "lidar_00000", "lidar_00001"]
