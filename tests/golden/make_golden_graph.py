"""Execute the reference's own ``air/transformer.py`` on the numpy TF shim (``tests/golden/tf_shim.py``) for the inputs of
the committed golden cases and store its outputs: ``tests/golden/graph_<case>.npz``.  Run from the repo root in the
authoring container (``python tests/golden/make_golden_graph.py``); the reference file is imported from where it lies."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

sys.modules["tensorflow"] = tf_shim
spec = importlib.util.spec_from_file_location("ref_transformer", "/root/reference/air/transformer.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = ["read_50_28", "write_28_50", "adversarial_17x23x3_9x31", "adversarial_50_28", "adversarial_28_50", "out_1x1", "out_1x7",
         "fullcover_64_28"]


def main():
    for name in CASES:
        z = np.load(os.path.join(HERE, name + ".npz"))
        U, theta = z["U"].astype(np.float32), z["theta"].astype(np.float32)
        out_size = tuple(int(v) for v in z["out_size"]) if "out_size" in z else tuple(z["out"].shape[1:3])
        finite = np.isfinite(theta).all(1) & (np.abs(theta).max(1) < 1e6)     # int32(floor(.)) is only defined in range
        out = np.asarray(ref.transformer(tf_shim._t(U[finite]), tf_shim._t(theta[finite]), out_size))
        np.savez_compressed(os.path.join(HERE, "graph_" + name + ".npz"), rows=np.nonzero(finite)[0], out=out.astype(np.float32))
        print(name, U.shape, "->", out.shape, "rows", int(finite.sum()), "of", len(finite))
    # batch_transformer: 3 images x 4 transforms
    rng = np.random.default_rng(5)
    U = rng.random((3, 20, 24, 2), dtype=np.float32)
    th = (np.array([0.5, 0, 0.1, 0, 0.6, -0.2], np.float32) + 0.3 * rng.normal(size=(3, 4, 6))).astype(np.float32)
    out = np.asarray(ref.batch_transformer(tf_shim._t(U), tf_shim._t(th), (9, 11)))
    np.savez_compressed(os.path.join(HERE, "graph_batch_transformer.npz"), U=U, thetas=th, out=out.astype(np.float32))
    print("batch_transformer", out.shape)


if __name__ == "__main__":
    main()
