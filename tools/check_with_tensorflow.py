"""For a maintainer who HAS TensorFlow 1.12 (Python 3.6; not installable in the authoring image, so this script has not
been run there): evaluate the reference's ``air/transformer.py`` with the real TensorFlow on the committed golden inputs
and compare with the committed outputs -- the one check that would turn the ``[TF-1.12 assumed]`` per-kernel numerics
(linspace recurrence, K = 3 accumulation order, ``add_n`` order) into facts.

    python tools/check_with_tensorflow.py /path/to/MOG-ASR        # the reference checkout

Prints, per golden case, whether the forward output is bit-identical and the largest absolute deviation, and the same for
the gradients w.r.t. U and theta (``tf.gradients``) against ``graph_grad_*.npz``."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = ["read_50_28", "write_28_50", "adversarial_17x23x3_9x31", "adversarial_50_28", "adversarial_28_50", "out_1x1", "out_1x7",
         "fullcover_64_28"]


def main(ref_root):
    sys.path.insert(0, ref_root)
    import tensorflow as tf                       # 1.12: graph mode
    from air.transformer import transformer
    for name in CASES:
        z = np.load(os.path.join(GOLD, name + ".npz"))
        g = np.load(os.path.join(GOLD, "graph_" + name + ".npz"))
        rows = g["rows"]
        tf.reset_default_graph()
        U = tf.placeholder(tf.float32, z["U"][rows].shape)
        th = tf.placeholder(tf.float32, z["theta"][rows].shape)
        out = transformer(U, th, [int(v) for v in z["out_size"]])
        feeds = {U: z["U"][rows], th: z["theta"][rows]}
        fetch = [out]
        gpath = os.path.join(GOLD, "graph_grad_" + name + ".npz")
        if os.path.exists(gpath):
            gout = tf.placeholder(tf.float32, z["gout"][rows].shape)
            fetch += tf.gradients(out, [U, th], grad_ys=gout)
            feeds[gout] = z["gout"][rows]
        with tf.Session(config=tf.ConfigProto(device_count={"GPU": 0})) as sess:
            res = sess.run(fetch, feeds)
        same = np.array_equal(res[0].view(np.uint32), g["out"].view(np.uint32))
        line = f"{name:28s} forward bit-identical: {same}  max|diff| {np.nanmax(np.abs(res[0] - g['out'])):.3g}"
        if len(res) == 3:
            gg = np.load(gpath)
            line += (f"  dU max|diff| {np.abs(res[1] - gg['dU']).max():.3g}"
                     f"  dtheta max|diff| {np.abs(res[2].reshape(-1, 2, 3) - gg['dtheta']).max():.3g}")
        print(line)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
