"""CPU: pins the oracle's geometry ops against everything the reference's own tests hold for them
(lmbspecialops/test/*.py, restated as pytest) plus hand-computed cases of the documented semantics."""
import os
import warnings

import numpy as np
import pytest

from oracle import ops as oops

TYPES = (np.float32, np.float64)


def angleaxis_to_rotation_matrix(aa):
    """lmbspecialops/test/helper.py:20-41 (Rodrigues)."""
    angle = np.sqrt(aa.dot(aa))
    if angle > 1e-6:
        c, s = np.cos(angle), np.sin(angle)
        u = aa / angle
        R = np.empty((3, 3))
        R[0, 0] = c + u[0] * u[0] * (1 - c); R[0, 1] = u[0] * u[1] * (1 - c) - u[2] * s; R[0, 2] = u[0] * u[2] * (1 - c) + u[1] * s
        R[1, 0] = u[1] * u[0] * (1 - c) + u[2] * s; R[1, 1] = c + u[1] * u[1] * (1 - c); R[1, 2] = u[1] * u[2] * (1 - c) - u[0] * s
        R[2, 0] = u[2] * u[0] * (1 - c) - u[1] * s; R[2, 1] = u[2] * u[1] * (1 - c) + u[0] * s; R[2, 2] = c + u[2] * u[2] * (1 - c)
        return R
    return np.eye(3)


def angleaxis_to_quaternion(aa):
    """lmbspecialops/test/helper.py:44-60."""
    angle = np.sqrt(aa.dot(aa))
    if angle > 1e-6:
        s = np.sin(0.5 * angle)
        return np.array([np.cos(0.5 * angle), s * aa[0] / angle, s * aa[1] / angle, s * aa[2] / angle])
    return np.array([1.0, 0, 0, 0])


# ---- median3x3_downsample: test_Median3x3Downsample.py:30-35,58-67 (exact) ----------------------
@pytest.mark.parametrize("dtype", TYPES)
def test_median_reference_kats(dtype, golden_dir):
    z = np.load(os.path.join(golden_dir, "reference_kats.npz"))
    A = z["median_in_single"].astype(dtype)
    assert np.array_equal(oops.median3x3_downsample(A), z["median_out_single"].astype(dtype))
    A = z["median_in_1d"].astype(dtype)
    assert np.array_equal(oops.median3x3_downsample(A), z["median_out_1d"].astype(dtype))
    assert np.array_equal(oops.median3x3_downsample(A.T.copy()), z["median_out_1d"].astype(dtype).T)


@pytest.mark.parametrize("dtype", TYPES)
def test_median_is_true_median_without_nans(dtype):
    rng = np.random.RandomState(0)
    A = rng.rand(3, 10, 13).astype(dtype)
    out = oops.median3x3_downsample(A)
    assert out.shape == (3, 5, 7)
    pad = np.pad(A, ((0, 0), (1, 1), (1, 1)), mode="edge")
    for y in range(5):
        for x in range(7):
            win = pad[:, 2 * y:2 * y + 3, 2 * x:2 * x + 3].reshape(3, 9)
            assert np.array_equal(out[:, y, x], np.sort(win, axis=1)[:, 4])


# ---- depth_to_flow o flow_to_depth2 round trip: test_FlowToDepth2.py:37-79 ----------------------
@pytest.mark.parametrize("dtype", TYPES)
@pytest.mark.parametrize("inverse_depth", (False, True))
@pytest.mark.parametrize("normalize_flow", (False, True))
def test_depth_flow_round_trip(dtype, inverse_depth, normalize_flow):
    rng = np.random.RandomState(int(inverse_depth) * 2 + int(normalize_flow))
    depth = rng.uniform(5, 10, (1, 1, 6, 12)).astype(dtype)
    if inverse_depth:
        depth = (1 / depth).astype(dtype)
    rotation = rng.uniform(0.0, 0.05, (1, 3)).astype(dtype)
    translation = (np.array([[1, 0, 0]]) + rng.uniform(-0.2, 0.2, (1, 3))).astype(dtype)
    intrinsics = np.array([[1, 1, 0.5, 0.5]]).astype(dtype)
    flow = oops.depth_to_flow(depth, intrinsics, rotation, translation, inverse_depth=inverse_depth,
                              normalize_flow=normalize_flow)
    assert flow.shape == (1, 2, 6, 12)
    q = angleaxis_to_quaternion(rotation[0].astype(np.float64))[None].astype(dtype)
    computed = oops.flow_to_depth2(flow, intrinsics, q, translation, inverse_depth=inverse_depth,
                                   normalized_flow=normalize_flow, rotation_format="quaternion")
    np.testing.assert_allclose(depth, computed, rtol=1e-4, atol=1e-4)


# ---- rotation formats agree: test_FlowToDepth2.py:83-133 ----------------------------------------
@pytest.mark.parametrize("dtype", TYPES)
def test_rotation_formats(dtype):
    rng = np.random.RandomState(5)
    depth = rng.uniform(5, 10, (1, 1, 6, 12)).astype(dtype)
    rotation = rng.uniform(0.0, 0.05, (1, 3)).astype(dtype)
    translation = (np.array([[1, 0, 0]]) + rng.uniform(-0.2, 0.2, (1, 3))).astype(dtype)
    intrinsics = np.array([[1, 1, 0.5, 0.5]]).astype(dtype)
    flow = oops.depth_to_flow(depth, intrinsics, rotation, translation)
    R = angleaxis_to_rotation_matrix(rotation[0].astype(np.float64))[None].astype(dtype)
    q = angleaxis_to_quaternion(rotation[0].astype(np.float64))[None].astype(dtype)
    d_aa = oops.flow_to_depth2(flow, intrinsics, rotation, translation, rotation_format="angleaxis3")
    d_R = oops.flow_to_depth2(flow, intrinsics, R, translation, rotation_format="matrix")
    d_q = oops.flow_to_depth2(flow, intrinsics, q, translation, rotation_format="quaternion")
    np.testing.assert_allclose(d_aa, d_R, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(depth, d_q, rtol=1e-4, atol=1e-4)
    # the oracle's Rodrigues is the published closed form (helpers.py:44-57)
    np.testing.assert_allclose(oops.rotation_matrix(rotation)[0], R[0], atol=1e-6 if dtype == np.float32 else 1e-14)


def test_flow_to_depth_deprecated_twin_warns_and_matches():
    rng = np.random.RandomState(1)
    flow = rng.uniform(-0.05, 0.05, (2, 2, 6, 8)).astype(np.float32)
    K = np.tile(np.array([[0.89115971, 1.18821287, 0.5, 0.5]], np.float32), (2, 1))
    r = rng.uniform(-0.05, 0.05, (2, 3)).astype(np.float32)
    t = (np.array([[1, 0, 0]]) + rng.uniform(-0.2, 0.2, (2, 3))).astype(np.float32)
    with pytest.warns(DeprecationWarning):
        a = oops.flow_to_depth(flow, K, r, t, normalized_flow=True, inverse_depth=True)
    b = oops.flow_to_depth2(flow, K, r, t, normalized_flow=True, inverse_depth=True)
    assert np.array_equal(a, b)


def test_depth_to_flow_invalid_depth_is_nan_and_flow_to_depth_invalid_is_zero():
    depth = np.array([[[[1.0, 0.0, -2.0, np.inf, np.nan, 0.5]]]], np.float32)
    K = np.array([[1, 1, 0.5, 0.5]], np.float32)
    flow = oops.depth_to_flow(depth, K, np.zeros((1, 3), np.float32), np.array([[1, 0, 0]], np.float32))
    assert np.isfinite(flow[0, :, 0, 0]).all() and np.isfinite(flow[0, :, 0, 5]).all()
    assert np.isnan(flow[0, :, 0, 1:5]).all()       # depthtoflow.cc:289-303
    # inverse depth: 1/inf = 0 is invalid, 1/0 = inf is invalid
    flow = oops.depth_to_flow(depth, K, np.zeros((1, 3), np.float32), np.array([[1, 0, 0]], np.float32), inverse_depth=True)
    assert np.isnan(flow[0, :, 0, 1:5]).all()
    # behind the camera / NaN flow -> 0 (flowtodepth.cc:464-474)
    f = np.zeros((1, 2, 1, 3), np.float32)
    f[0, 0, 0, 0] = 0.1      # camera moved to +x => points move to -x; positive flow triangulates behind the camera
    f[0, 0, 0, 1] = np.nan
    f[0, 0, 0, 2] = -0.1
    d = oops.flow_to_depth2(f, K, np.zeros((1, 3), np.float32), np.array([[-1, 0, 0]], np.float32), normalized_flow=True)
    assert d[0, 0, 0, 1] == 0
    assert (d[0, 0, 0, 0] == 0) != (d[0, 0, 0, 2] == 0)


# ---- shapes: test_FlowToDepth2.py:145-200, test_ScaleInvariantGradient.py:65-77 ------------------
def test_shapes():
    assert oops.flow_to_depth2(np.zeros((2, 6, 12), np.float32), np.array([1, 1, .5, .5], np.float32),
                               np.zeros(3, np.float32), np.array([1, 0, 0], np.float32)).shape == (1, 1, 6, 12)
    b = 7
    assert oops.flow_to_depth2(np.zeros((b, 2, 6, 12), np.float32), np.zeros((b, 4), np.float32),
                               np.zeros((b, 3), np.float32), np.zeros((b, 3), np.float32)).shape == (b, 1, 6, 12)
    for shape in ((8, 40, 31), (8, 1, 40, 31), (2, 2, 2, 40, 31)):
        assert oops.scale_invariant_gradient(np.ones(shape, np.float32)).shape == (8, 2, 40, 31)
    batch = np.array([7, 7, 7, 5])
    for i in range(4):
        batch = np.roll(batch, 1)
        with pytest.raises(ValueError, match="Dimensions must be equal"):
            oops.flow_to_depth2(np.zeros((batch[0], 2, 6, 12), np.float32), np.zeros((batch[3], 4), np.float32),
                                np.zeros((batch[1], 3), np.float32), np.zeros((batch[2], 3), np.float32))


# ---- scale_invariant_gradient: the formula of the op doc (scaleinvariantgradient.cc:48-70) --------
@pytest.mark.parametrize("dtype", TYPES)
def test_sig_matches_formula(dtype):
    A = np.linspace(1, 2, num=16, dtype=dtype).reshape(4, 4)   # the input of test_ScaleInvariantGradient.py:33
    deltas, weights, eps = [1, 2, 4], [1, 0.5, 0.25], 0.001
    out = oops.scale_invariant_gradient(A, deltas, weights, eps)
    assert out.shape == (1, 2, 4, 4)
    ref = np.zeros((2, 4, 4), np.float64)
    for y in range(4):
        for x in range(4):
            for d, w in zip(deltas, weights):
                vx = A[y, x + d] if 0 <= x + d < 4 else A[y, x]
                vy = A[y + d, x] if 0 <= y + d < 4 else A[y, x]
                ref[0, y, x] += w * (float(vx) - float(A[y, x])) / (abs(float(A[y, x])) + abs(float(vx)) + eps)
                ref[1, y, x] += w * (float(vy) - float(A[y, x])) / (abs(float(A[y, x])) + abs(float(vy)) + eps)
    np.testing.assert_allclose(out[0], ref, rtol=2e-6 if dtype == np.float32 else 1e-9, atol=1e-7)


def test_sig_numeric_gradient_property():
    """test_ScaleInvariantGradient.py:31-46 checks d(out)/d(in) numerically; the forward-only equivalent:
    scaling the input by a constant leaves the output (almost, eps) unchanged -- the op's namesake property."""
    rng = np.random.RandomState(3)
    A = rng.uniform(1, 2, (1, 12, 9))
    a = oops.scale_invariant_gradient(A, [1, 2], [1, 0.5], 1e-9)
    b = oops.scale_invariant_gradient(A * 37.0, [1, 2], [1, 0.5], 1e-9)
    np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("dtype", TYPES)
def test_leaky_relu(dtype):
    x = np.array([-2, -0.5, 0, 0.5, 3], dtype)
    np.testing.assert_allclose(oops.leaky_relu(x, 0.2), np.maximum(np.float32(0.2).astype(dtype) * x, x), rtol=0, atol=0)
    np.testing.assert_allclose(oops.leaky_relu(x), np.maximum(dtype(np.float32(0.1)) * x, x), rtol=0, atol=0)


# ---- warp2d: no reference test exists (parity unpinned by the reference); semantics of warp2d.cc ---
@pytest.mark.parametrize("dtype", TYPES)
def test_warp2d_hand_cases(dtype):
    img = np.arange(2 * 4 * 5, dtype=dtype).reshape(1, 2, 4, 5)
    zero = np.zeros((1, 2, 4, 5), dtype)
    # zero displacement: interior pixels unchanged; VALUE mode kills the last row/column because x3 = x0+1
    # must be < W (warp2d.cc:236), CLAMP mode keeps them
    out_c = oops.warp2d(img, zero, border_mode="clamp")
    assert np.array_equal(out_c, img)
    out_v = oops.warp2d(img, zero, border_mode="value", border_value=-7)
    assert np.array_equal(out_v[..., :3, :4], img[..., :3, :4])
    assert (out_v[..., 3, :] == -7).all() and (out_v[..., :, 4] == -7).all()
    # +1 pixel in x, unnormalized and normalized
    d = zero.copy(); d[:, 0] = 1
    out = oops.warp2d(img, d, border_mode="clamp")
    assert np.array_equal(out[..., :, :4], img[..., :, 1:])
    dn = zero.copy(); dn[:, 0] = 1.0 / 5
    np.testing.assert_allclose(oops.warp2d(img, dn, normalized=True, border_mode="clamp"), out, atol=1e-5)
    # half pixel: bilinear average
    d = zero.copy(); d[:, 0] = 0.5
    out = oops.warp2d(img, d, border_mode="clamp")
    np.testing.assert_allclose(out[..., :, :4], (img[..., :, :4] + img[..., :, 1:]) / 2)
    # truncation toward zero: p2.x in (-1,0) has p2i = 0 and a NEGATIVE weight a (extrapolation), and is
    # "valid" in VALUE mode
    d = zero.copy(); d[:, 0] = -0.25
    out = oops.warp2d(img, d, border_mode="value", border_value=-7)
    np.testing.assert_allclose(out[0, 0, 0, 0], img[0, 0, 0, 0] * 1.25 - 0.25 * img[0, 0, 0, 1])
    # NaN / huge displacement: x86 cvttss2si gives INT_MIN -> border value in VALUE mode
    d = zero.copy(); d[0, 0, 1, 1] = np.nan; d[0, 1, 2, 2] = 1e20
    out = oops.warp2d(img, d, border_mode="value", border_value=-7)
    assert out[0, 0, 1, 1] == -7 and out[0, 1, 2, 2] == -7
    # rank 2 and rank 3 inputs (warp2d.cc:150-160)
    assert oops.warp2d(img[0, 0], zero[0], border_mode="clamp").shape == (4, 5)
    assert oops.warp2d(img[0], zero[0], border_mode="clamp").shape == (2, 4, 5)


def test_sculpture_ground_truth_semantics(sculpture):
    """Semantic pin on real data (examples/create_dataset_and_use_readerop.py:16,24-37): the flow that
    depth_to_flow derives from the ground-truth depth and pose of the example pair must make
    warp2d(image2) look like image1 (photometric error drops by a large factor)."""
    K = np.array([[0.89115971, 1.18821287, 0.5, 0.5]], np.float32)
    Rt2 = sculpture["Rt2"]
    R, t = Rt2[:, :3], Rt2[:, 3]
    depth = sculpture["depth1_l2"][None, None].astype(np.float32)   # 48x64
    flow = oops.depth_to_flow(depth, K, R[None].astype(np.float32), t[None].astype(np.float32), rotation_format="matrix",
                              normalize_flow=True)
    valid = depth[0, 0] > 0                      # the ground truth has holes (depth 0 -> NaN flow, depthtoflow.cc:289)
    assert np.isnan(flow[0, :, ~valid]).all() and np.isfinite(flow[0, :, valid]).all()
    img1 = sculpture["image1"][:, :, 2::4, 2::4]
    img2 = sculpture["image_pair"][:, 3:6, 2::4, 2::4]
    flow0 = np.where(np.isnan(flow), 0, flow).astype(np.float32)
    warped = oops.warp2d(np.ascontiguousarray(img2), flow0, normalized=True, border_mode="clamp")
    inside = valid & (np.abs(flow0[0]).max(axis=0) < 0.5)
    before = np.abs(img2 - img1)[0][:, inside].mean()
    after = np.abs(warped - img1)[0][:, inside].mean()
    assert after < 0.75 * before, (before, after)   # point-sampled 48x64 images, occlusions included
    # and flow_to_depth inverts it
    d = oops.flow_to_depth2(flow0, K, R[None].astype(np.float32), t[None].astype(np.float32), rotation_format="matrix",
                            normalized_flow=True)
    np.testing.assert_allclose(d[0, 0][valid], depth[0, 0][valid], rtol=2e-3)
