"""Generates tests/golden/*.npz.  Run in the build container (needs /root/reference for the example
images); the GPU box only reads the committed .npz files.

  sculpture_inputs.npz   the reference's example pair prepared exactly like examples/example.py:15-42
                         (PIL resize with the reference-era NEAREST default for image2_2), stored as
                         uint8 so the fixture stays small, plus the ground truth of
                         examples/sculpture_depth1.npy / sculpture_Rt{1,2}.txt at 48x64.
  reference_kats.npz     the known-answer vectors of the reference's own op tests
                         (lmbspecialops/test/test_Median3x3Downsample.py:30-35,58-67).
  oracle_config1.npz     BASELINE.json configs[0]: netFlow1 on the sculpture pair, CPU oracle fp32,
                         synthetic weights seed 0 -> predict_flowconf2.  ORACLE-GENERATED (the reference
                         has no network-level golden output; parity unpinned at that level).
  oracle_pipeline.npz    full pipeline on the same pair (fp32 and fp64 oracle): depth0 at half precision
                         of storage is NOT used -- stored as float32.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/examples"


def main():
    from PIL import Image
    import torch
    from demon_b200 import weights as W
    from oracle.network import OracleNets, flow_block

    img1 = Image.open(os.path.join(REF, "sculpture1.png")).convert("RGB")
    img2 = Image.open(os.path.join(REF, "sculpture2.png")).convert("RGB")
    assert img1.size == (256, 192) and img2.size == (256, 192)
    img2_2 = img2.resize((64, 48), Image.NEAREST)   # Pillow 2.0's default filter (Dockerfile:15)
    a1, a2, a22 = np.array(img1), np.array(img2), np.array(img2_2)
    depth1 = np.load(os.path.join(REF, "sculpture_depth1.npy")).astype(np.float32)
    Rt1 = np.loadtxt(os.path.join(REF, "sculpture_Rt1.txt")).astype(np.float64)
    Rt2 = np.loadtxt(os.path.join(REF, "sculpture_Rt2.txt")).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "sculpture_inputs.npz"), img1=a1, img2=a2, img2_2=a22,
                        depth1_l2=depth1[2::4, 2::4].copy(), Rt1=Rt1, Rt2=Rt2)

    np.savez(os.path.join(HERE, "reference_kats.npz"),
             median_in_1d=np.array([[1, 2, 3, 4, 5]], np.float32), median_out_1d=np.array([[1, 3, 5]], np.float32),
             median_in_single=np.array([[1]], np.float32), median_out_single=np.array([[1]], np.float32))

    def prep(a):
        return (a.astype(np.float32) / 255 - 0.5).transpose(2, 0, 1)
    image_pair = np.concatenate((prep(a1), prep(a2)), axis=0)[None]
    image2_2 = prep(a22)[None]
    w = W.synthetic_weights(0)
    net = OracleNets(w)
    with torch.no_grad():
        f = flow_block(net.W, "netFlow1", torch.from_numpy(image_pair))
    np.savez_compressed(os.path.join(HERE, "oracle_config1.npz"),
                        predict_flowconf2=f["predict_flowconf2"].numpy(), predict_flowconf5=f["predict_flowconf5"].numpy())
    r32 = net.pipeline(image_pair, image2_2)
    r64 = OracleNets(w, torch.float64).pipeline(image_pair, image2_2)
    keys = ["predict_depth0", "predict_flow2", "predict_depth2", "predict_normal2", "predict_rotation", "predict_translation"]
    out = {k + "_f32": r32[k].numpy() for k in keys}
    out.update({k + "_f64": r64[k].numpy().astype(np.float32) for k in keys})
    np.savez_compressed(os.path.join(HERE, "oracle_pipeline.npz"), **out)
    d32, d64 = r32["predict_depth0"].numpy(), r64["predict_depth0"].numpy()
    print("fp32 oracle vs fp64 oracle: depth0 L1-rel %.3e" % (np.abs(d32 - d64).sum() / np.abs(d64).sum()))
    f32, f64 = r32["predict_flow2"].numpy(), r64["predict_flow2"].numpy()
    print("fp32 oracle vs fp64 oracle: flow2 EPE %.3e" % np.sqrt(((f32 - f64) ** 2).sum(1)).mean())


if __name__ == "__main__":
    main()
