"""Golden vectors for InfiniteSceneGeneration.prepare_pcd (sgam/inference_pipeline.py:1014-1036), produced by the
UNMODIFIED reference method.  Runs only in the build container (needs /root/reference).

    python tests/golden/make_golden_pcd.py

prepare_pcd wraps its result in an Open3D PointCloud; Open3D is absent, so `o3d` is stubbed with a record that keeps
the arrays it is given (pure glue: the arithmetic under test is the numpy code of the reference method)."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    out_path = os.path.join(HERE, "prepare_pcd_vectors.npz")
    mg.install_shims()
    import sgam.inference_pipeline as ip
    o3d = types.SimpleNamespace(
        geometry=types.SimpleNamespace(PointCloud=lambda: types.SimpleNamespace(points=None, colors=None)),
        utility=types.SimpleNamespace(Vector3dVector=lambda a: np.array(a)))
    ip.o3d = o3d
    pipe = object.__new__(ip.InfiniteSceneGeneration)
    rng = np.random.default_rng(7)
    out = {}
    for name, (H, W, K) in {"clevr": (24, 32, np.array([[355.5555, 0, 128], [0, 355.5555, 128], [0, 0, 1]])),
                            "ge": (16, 16, np.array([[248.88887, 0, 128], [0, 248.88887, 128], [0, 0, 1]]))}.items():
        depth = rng.uniform(1.0, 16.0, (H, W)).astype(np.float32)
        depth[0, :3] = 0.0
        color = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        c, s = np.cos(0.3), np.sin(0.3)
        c2w = np.eye(4)
        c2w[:3, :3] = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
        c2w[:3, 3] = rng.uniform(-20, 5, 3)
        Rt = np.linalg.inv(c2w @ np.diag([1., -1., -1., 1.]))
        pcd = pipe.prepare_pcd(depth, color, K, Rt)
        out.update({f"{name}_depth": depth, f"{name}_color": color, f"{name}_K": K, f"{name}_Rt": Rt,
                    f"{name}_points": np.asarray(pcd.points), f"{name}_colors": np.asarray(pcd.colors)})
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
