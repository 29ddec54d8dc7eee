"""Variable table of the DeMoN `networks_original` graphs and a seeded synthetic
weight generator.

Variable names and layouts are the ones TensorFlow creates for the reference
graph, so that a reader of the reference checkpoint (`weights/demon_original`,
examples/example.py:82-83) can feed this package unchanged:

  * `tf.layers.conv2d(name=N)` under `tf.variable_scope(S)`  ->  `S/N/kernel`
    with layout [kh, kw, cin, cout] and `S/N/bias` [cout]
    (python/depthmotionnet/helpers.py:86-94,130-153);
  * `tf.layers.conv2d_transpose(name='upconv')`  ->  `S/.../upconv/kernel` with
    layout [kh, kw, cout, cin] (blocks_original.py:64-74,97-108);
  * `tf.layers.dense(name=N)` -> `S/N/kernel` [in, out] (blocks_original.py:390-410).

The pretrained checkpoint is not in the reference repository
(weights/download_weights.sh:2 is a wget) so tests and benchmarks use
`synthetic_weights`.
"""
from collections import OrderedDict

import numpy as np


def _sep(name, k, cin, cmid, cout):
    """convrelu2_caffe_padding: (k x 1) then (1 x k), helpers.py:105-153."""
    return [(name + "y", "conv", (k, 1, cin, cmid)), (name + "x", "conv", (1, k, cmid, cout))]


def _trunk(extra_in, conv2_out, conv5_k):
    specs = []
    specs += _sep("conv1", 9, 6, 32, 32)
    specs += _sep("conv2", 7, 32, conv2_out, conv2_out)
    if extra_in:
        specs += _sep("conv2_extra_inputs", 3, extra_in, 32, 32)
    specs += _sep("conv2_1", 3, 64, 64, 64)
    specs += _sep("conv3", 5, 64, 128, 128)
    specs += _sep("conv3_1", 3, 128, 128, 128)
    specs += _sep("conv4", 5, 128, 256, 256)
    specs += _sep("conv4_1", 3, 256, 256, 256)
    specs += _sep("conv5", conv5_k, 256, 512, 512)
    specs += _sep("conv5_1", 3, 512, 512, 512)
    return specs


def flow_block_specs(iterative):
    """flow_block_demon_original, blocks_original.py:121-235."""
    specs = _trunk(9 if iterative else 0, 32 if iterative else 64, 5)
    specs += [
        ("predict_flow5/conv1", "conv", (3, 3, 512, 24)),
        ("predict_flow5/conv2", "conv", (3, 3, 24, 4)),
        ("upsample_flow5to4/upconv", "deconv", (4, 4, 2, 4)),
        ("refine4/upconv", "deconv", (4, 4, 256, 512)),
        ("refine3/upconv", "deconv", (4, 4, 128, 514)),
        ("refine2/upconv", "deconv", (4, 4, 64, 256)),
        ("predict_flow2/conv1", "conv", (3, 3, 128, 24)),
        ("predict_flow2/conv2", "conv", (3, 3, 24, 4)),
    ]
    return specs


def depthmotion_block_specs(iterative):
    """depthmotion_block_demon_original, blocks_original.py:299-448."""
    specs = _trunk(8 if iterative else 7, 32, 3)
    specs += [
        ("motion_conv1", "conv", (3, 3, 512, 128)),
        ("motion_fc1", "dense", (6144, 1024)),
        ("motion_fc2", "dense", (1024, 128)),
        ("motion_fc3", "dense", (128, 7)),
        ("refine4/upconv", "deconv", (4, 4, 256, 512)),
        ("refine3/upconv", "deconv", (4, 4, 128, 512)),
        ("refine2/upconv", "deconv", (4, 4, 64, 256)),
        ("predict_depthnormal2/conv1", "conv", (3, 3, 128, 24)),
        ("predict_depthnormal2/conv2", "conv", (3, 3, 24, 4)),
    ]
    return specs


def refine_block_specs():
    """depth_refine_block_demon_original, blocks_original.py:452-513."""
    return [
        ("conv0", "conv", (3, 3, 4, 32)),
        ("conv1", "conv", (3, 3, 32, 64)),
        ("conv1_1", "conv", (3, 3, 64, 64)),
        ("conv2", "conv", (3, 3, 64, 128)),
        ("conv2_1", "conv", (3, 3, 128, 128)),
        ("refine1/upconv", "deconv", (4, 4, 64, 128)),
        ("refine0/upconv", "deconv", (4, 4, 32, 128)),
        ("predict_depth0/conv1", "conv", (3, 3, 64, 16)),
        ("predict_depth0/conv2", "conv", (3, 3, 16, 1)),
    ]


# scope -> block specs; scopes as in networks_original.py:44,50,125,142,227
SCOPES = OrderedDict([
    ("netFlow1", lambda: flow_block_specs(False)),
    ("netDM1", lambda: depthmotion_block_specs(False)),
    ("netFlow2", lambda: flow_block_specs(True)),
    ("netDM2", lambda: depthmotion_block_specs(True)),
    ("netRefine", refine_block_specs),
])


def variable_specs():
    """OrderedDict: full variable name -> (kind, shape) for kernels and biases."""
    out = OrderedDict()
    for scope, fn in SCOPES.items():
        for name, kind, shape in fn():
            out["%s/%s/kernel" % (scope, name)] = (kind, tuple(shape))
            nout = shape[2] if kind == "deconv" else shape[-1]
            out["%s/%s/bias" % (scope, name)] = ("bias", (nout,))
    return out


_TRUNK_RES = {"conv1y": (96, 256), "conv1x": (96, 128), "conv2y": (48, 128), "conv2x": (48, 64),
              "conv2_extra_inputsy": (48, 64), "conv2_extra_inputsx": (48, 64),
              "conv2_1y": (48, 64), "conv2_1x": (48, 64),
              "conv3y": (24, 64), "conv3x": (24, 32), "conv3_1y": (24, 32), "conv3_1x": (24, 32),
              "conv4y": (12, 32), "conv4x": (12, 16), "conv4_1y": (12, 16), "conv4_1x": (12, 16),
              "conv5y": (6, 16), "conv5x": (6, 8), "conv5_1y": (6, 8), "conv5_1x": (6, 8),
              "predict_flow5/conv1": (6, 8), "predict_flow5/conv2": (6, 8), "motion_conv1": (6, 8),
              "upsample_flow5to4/upconv": (6, 8), "refine4/upconv": (6, 8), "refine3/upconv": (12, 16),
              "refine2/upconv": (24, 32),
              "predict_flow2/conv1": (48, 64), "predict_flow2/conv2": (48, 64),
              "predict_depthnormal2/conv1": (48, 64), "predict_depthnormal2/conv2": (48, 64)}


def layer_macs(refine_hw=(192, 256)):
    """OrderedDict: full layer name ("netFlow1/conv1y") -> algorithmic multiply-accumulates per image pair
    and per call of that layer.  conv: output pixels x kh x kw x cin x cout; transposed conv: input pixels x
    16 taps x cin x cout (every input pixel meets every tap once); dense: in x out.  No padding waste, no
    precision-split multiplier."""
    h, w = refine_hw
    refine_res = {"conv0": (h, w), "conv1": (h // 2, w // 2), "conv1_1": (h // 2, w // 2), "conv2": (h // 4, w // 4),
                  "conv2_1": (h // 4, w // 4), "refine1/upconv": (h // 4, w // 4), "refine0/upconv": (h // 2, w // 2),
                  "predict_depth0/conv1": (h, w), "predict_depth0/conv2": (h, w)}
    out = OrderedDict()
    for scope, fn in SCOPES.items():
        res = refine_res if scope == "netRefine" else _TRUNK_RES
        for name, kind, shape in fn():
            if kind == "dense":
                out[scope + "/" + name] = shape[0] * shape[1]
            else:   # for the transposed convs `res` is the INPUT resolution, for convs the OUTPUT resolution
                rh, rw = res[name]
                out[scope + "/" + name] = rh * rw * shape[0] * shape[1] * shape[2] * shape[3]
    return out


def macs_per_pair():
    """Algorithmic multiply-accumulates of one image pair through the full pipeline (bootstrap + 3 x iterative +
    refinement) at 256x192: 15 176.3 M = 30.353 GFLOP."""
    lm = layer_macs()
    tot = {scope: sum(v for k, v in lm.items() if k.startswith(scope + "/")) for scope in SCOPES}
    tot["pipeline"] = tot["netFlow1"] + tot["netDM1"] + 3 * (tot["netFlow2"] + tot["netDM2"]) + tot["netRefine"]
    tot["refine_fn"] = lambda h, w: sum(v for k, v in layer_macs((h, w)).items() if k.startswith("netRefine/"))
    return tot


def synthetic_weights(seed=0, dtype=np.float32):
    """Seeded stand-in for the (absent) pretrained checkpoint.

    Kernels: variance-scaling fan-in normal, stddev sqrt(2/fan_in)
    (helpers.py:66-67 `variance_scaling_initializer()` defaults).  Biases are
    small and NON-zero so the bias path is exercised.  The three prediction
    heads are re-scaled so that the geometry ops between the blocks see
    a geometrically SELF-CONSISTENT regime, like a trained DeMoN does: inverse
    depth around 0.5, rotation of a few hundredths of a radian, translation near
    (0.9, 0.1, -0.05), scale near 1, and a flow prediction scattered around the
    flow that this depth and motion imply, depth_to_flow(0.5, r, t) ~ (0.381,
    0.035) of the image size.  Self-consistency matters for testing: where the
    predicted flow contradicts the predicted motion the triangulation of
    flow_to_depth (blocks_original.py:344) is singular (depth -> 1/0) and the
    reference network itself amplifies one float ulp of its own flow
    prediction into percent-level changes of its output -- no two float
    implementations, the reference's CPU and GPU paths included, agree there.
    """
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, (kind, shape) in variable_specs().items():
        if kind == "bias":
            out[name] = rng.uniform(-0.05, 0.05, size=shape).astype(dtype)
            continue
        if kind == "dense":
            fan_in = shape[0]
        elif kind == "deconv":  # each output pixel meets 4 of the 16 taps of every input channel
            fan_in = shape[3] * 4
        else:
            fan_in = shape[0] * shape[1] * shape[2]
        out[name] = (rng.standard_normal(size=shape) * np.sqrt(2.0 / fan_in)).astype(dtype)

    def rescale(prefix, wscale, bias):
        out[prefix + "/kernel"] = (out[prefix + "/kernel"] * wscale).astype(dtype)
        out[prefix + "/bias"] = np.asarray(bias, dtype=dtype)

    for scope in ("netFlow1", "netFlow2"):
        rescale(scope + "/predict_flow5/conv2", 0.02, [0.381, 0.035, 0.3, 0.3])
        rescale(scope + "/predict_flow2/conv2", 0.02, [0.381, 0.035, 0.3, 0.3])
    for scope in ("netDM1", "netDM2"):
        rescale(scope + "/predict_depthnormal2/conv2", 0.1, [0.5, 0.0, 0.0, -0.8])
        rescale(scope + "/motion_fc3", 0.05, [0.02, -0.03, 0.01, 0.9, 0.1, -0.05, 1.0])
    rescale("netRefine/predict_depth0/conv2", 0.2, [0.5])
    return out
