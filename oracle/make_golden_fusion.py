"""Record tests/golden/fusion_m9.npz by RUNNING THE REFERENCE's multiview_fusion (v0520.py:456-484, with its own
ScaledDotProductAttention / VisualProjectionHeadPretrain, utils_v0511.py) through oracle/ref_shim.py.  Container only.

    python oracle/make_golden_fusion.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
import test_fusion as tf  # noqa: E402


def main():
    d, d_out, p, b = 32, 16, 5, 5
    ref = ref_shim.make_fusion_self(d, d_out, seed=7)
    ref.train(True)                                    # BatchNorm in training mode, as in the pre-training step
    ref.multiview_cross_attention.dropout.p = 0.0
    sd = {f"sd.{k}": v.clone().numpy() for k, v in ref.state_dict().items()}      # before the BN statistics move
    gi, li = tf._inputs(len(tf.IDS), p, d, seed=2)
    g = gi.clone().requires_grad_(True)
    l = li.clone().requires_grad_(True)
    o0, o1 = ref_shim.multiview_fusion(ref, g, l, tf.IDS, b)
    (o0.square().sum() + (o1 * 0.3).sum()).backward()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fusion_m9.npz"), d=d, d_out=d_out, p=p, b=b, ids=tf.IDS,
                        **{"global": gi.numpy(), "local": li.numpy()}, out_global=o0.detach().numpy(),
                        out_local=o1.detach().numpy(), d_global=g.grad.numpy(), d_local=l.grad.numpy(),
                        d_fc_k=ref.multiview_cross_attention.fc_k.weight.grad.numpy(), **sd)
    print("wrote fusion_m9.npz", float(o0.abs().mean()))


if __name__ == "__main__":
    main()
