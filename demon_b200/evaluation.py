"""Mirror of `depthmotionnet.evaluation.metrics` (python/depthmotionnet/evaluation/metrics.py) over the device
reductions in libdemon_b200.so (csrc/metrics.cu): same function names, arguments and result dictionaries, numpy or torch
arrays in, Python floats out -- plus batched forms that keep everything on the GPU.

    errs, errs_scaled = evaluate_depth(translation_gt, depth_gt, depth_pred)        # metrics.py:321-372
    errs = compute_errors(depth_pred, depth_gt)                                    # metrics.py:240-280
    epe = compute_flow_epe(flow_pred, flow_gt)                                     # metrics.py:377-387
    rot_deg, t_dist, t_deg = compute_motion_errors(pred6, gt6, True)               # metrics.py:390-445 (host, 6 numbers)

One streaming pass over prediction and ground truth yields all sums the eleven distances and the least-squares scale
factor need; the scale factor itself is computed on the device, so `evaluate_depth` is two kernel passes and one
[n,16]-double copy.  Tolerance against the numpy reference: 1e-5 relative (float32 pairwise summation there, double
accumulation here; `logf` vs numpy's log); counts (`num_valid`, the ratio thresholds) can differ by pixels whose
log-ratio sits within an ulp of the threshold.  There is no CPU fallback.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib

DISTANCES = ['l1', 'l1_inverse', 'scale_invariant', 'abs_relative', 'sq_relative', 'avg_log10', 'rmse_log', 'rmse',
             'ratio_threshold_1.25', 'ratio_threshold_1.5625', 'ratio_threshold_1.953125']
_SCALING = {'abs': 0, 'log': 1, 'inv': 2}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(x):
    if not torch.cuda.is_available():
        raise RuntimeError("demon_b200.evaluation needs a CUDA device (there is no CPU fallback)")
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32))
    return t.to(device="cuda", dtype=torch.float32).contiguous()


def depth_error_sums(pred, gt, inverse_pred=False, inverse_gt=False, gt_div=None, pred_scale=None):
    """pred, gt: [n, ...] (all trailing dims are pixels) -> CUDA float64 tensor [n, 16] of masked sums
    (include/demon_b200.h: demon_depth_error_sums_f32)."""
    p, g = _dev(pred), _dev(gt)
    if p.shape != g.shape:
        raise ValueError("prediction %s and ground truth %s differ in shape" % (tuple(p.shape), tuple(g.shape)))
    n = p.shape[0]
    hw = p[0].numel() if n else 0
    lib = _lib.load()
    sums = torch.empty((n, 16), dtype=torch.float64, device=p.device)
    ws = torch.empty(max(1, lib.demon_metric_workspace_bytes(n, hw) // 8), dtype=torch.float64, device=p.device)
    gd = None if gt_div is None else _dev(gt_div).reshape(n)
    ps = None if pred_scale is None else _dev(pred_scale).reshape(n)
    _lib.check(lib.demon_depth_error_sums_f32(p.data_ptr(), g.data_ptr(), n, hw, int(bool(inverse_pred)), int(bool(inverse_gt)),
                                              None if gd is None else gd.data_ptr(), None if ps is None else ps.data_ptr(),
                                              sums.data_ptr(), ws.data_ptr(), _stream()))
    return sums


def depth_scale_factor(sums, depth_scaling='abs'):
    """CUDA float32 [n]: the factor for the prediction that minimises the squared error (metrics.py:283-318)."""
    if depth_scaling not in _SCALING:
        raise Exception('Unknown depth scaling method')
    n = sums.shape[0]
    scale = torch.empty(n, dtype=torch.float32, device=sums.device)
    _lib.check(_lib.load().demon_depth_scale_factor(sums.data_ptr(), n, _SCALING[depth_scaling], scale.data_ptr(), _stream()))
    return scale


def errors_from_sums(row, distances_to_compute=None):
    """One row of sums (16 numbers on the host) -> the result dictionary of compute_errors (metrics.py:240-280)."""
    s = [float(v) for v in row]
    num = s[0]
    nan = float('nan')

    def dist(name):
        if num == 0:
            return nan
        if name == 'l1':
            return s[1] / num
        if name == 'l1_inverse':
            return s[2] / num
        if name == 'scale_invariant':
            return math.sqrt(max(0.0, s[4] / num - (s[3] * s[3]) / (num * num)))
        if name == 'abs_relative':
            return s[5] / num
        if name == 'sq_relative':
            return s[6] / num
        if name == 'avg_log10':
            return s[7] / num
        if name == 'rmse_log':
            return math.sqrt(s[4] / num)
        if name == 'rmse':
            return math.sqrt(s[8] / num)
        if name.startswith('ratio_threshold'):
            t = float(name.split('_')[-1])
            idx = {1.25: 9, 1.5625: 10, 1.953125: 11}.get(t)
            if idx is None:
                raise ValueError("ratio thresholds on the device are 1.25, 1.5625 and 1.953125 (got %r)" % t)
            return s[idx] / num
        raise KeyError(name)
    out = {'num_valid': int(num)}
    for name in (DISTANCES if distances_to_compute is None else distances_to_compute):
        out[name] = dist(name)
    return out


def compute_errors(depth_pred, depth_gt, distances_to_compute=None):
    """metrics.py:240-280 for one pair of depth maps (any shape)."""
    p, g = _dev(depth_pred).reshape(1, -1), _dev(depth_gt).reshape(1, -1)
    return errors_from_sums(depth_error_sums(p, g)[0].cpu().numpy(), distances_to_compute)


def evaluate_depth_batch(translation_gt, depth_gt_in, depth_pred_in, inverse_gt=True, inverse_pred=True, depth_scaling='abs'):
    """Batched evaluate_depth on the device: translation_gt [n,3], depths [n,...] -> (sums, sums_scaled, scale), CUDA
    tensors [n,16] / [n,16] / [n]; nothing leaves the GPU."""
    t = np.asarray(translation_gt.detach().cpu() if isinstance(translation_gt, torch.Tensor) else translation_gt, dtype=np.float64).reshape(-1, 3)
    norm = np.sqrt((t * t).sum(axis=1))
    gt_div = None if np.all(np.isclose(1.0, norm)) else np.where(np.isclose(1.0, norm), 1.0, norm).astype(np.float32)
    sums = depth_error_sums(depth_pred_in, depth_gt_in, inverse_pred, inverse_gt, gt_div)
    scale = depth_scale_factor(sums, depth_scaling)
    sums_scaled = depth_error_sums(depth_pred_in, depth_gt_in, inverse_pred, inverse_gt, gt_div, scale)
    return sums, sums_scaled, scale


def evaluate_depth(translation_gt, depth_gt_in, depth_pred_in, distances_to_compute=None, inverse_gt=True, inverse_pred=True,
                   depth_scaling='abs', depth_pred_max=np.inf):
    """metrics.py:321-372: (errs, errs_pred_scaled) for one sample."""
    p, g = _dev(depth_pred_in).reshape(1, -1), _dev(depth_gt_in).reshape(1, -1)
    sums, sums_scaled, _ = evaluate_depth_batch(np.asarray(translation_gt, dtype=np.float64).reshape(1, 3), g, p, inverse_gt, inverse_pred, depth_scaling)
    both = torch.stack([sums[0], sums_scaled[0]]).cpu().numpy()
    return errors_from_sums(both[0], distances_to_compute), errors_from_sums(both[1], distances_to_compute)


def flow_epe_sums(flow1, flow2):
    """flow [n,2,...] -> CUDA float64 [n,2] = (sum of the valid end point errors, count)."""
    a, b = _dev(flow1), _dev(flow2)
    if a.shape != b.shape or a.dim() < 3 or a.shape[1] != 2:
        raise ValueError("flows must be [n,2,...] of equal shape")
    n = a.shape[0]
    hw = a[0, 0].numel() if n else 0
    lib = _lib.load()
    sums = torch.empty((n, 2), dtype=torch.float64, device=a.device)
    ws = torch.empty(max(1, lib.demon_metric_workspace_bytes(n, hw) // 8), dtype=torch.float64, device=a.device)
    _lib.check(lib.demon_flow_epe_sums_f32(a.data_ptr(), b.data_ptr(), n, hw, sums.data_ptr(), ws.data_ptr(), _stream()))
    return sums


def compute_flow_epe(flow1, flow2):
    """metrics.py:377-387: average end point error between two flow fields [2,h,w]."""
    a, b = _dev(flow1), _dev(flow2)
    s = flow_epe_sums(a.reshape((1,) + tuple(a.shape)), b.reshape((1,) + tuple(b.shape)))[0].cpu().numpy()
    return float(s[0] / s[1]) if s[1] > 0 else float('nan')


def compute_motion_errors(predicted_motion, gt_motion, normalize_translations):
    """metrics.py:390-445 (six numbers per sample: host arithmetic in float64; minieigen's Quaternion(angle, axis) and
    angularDistance restated: 2 * acos(min(1, |q1 . q2|)))."""
    def quat(aa):
        aa = np.asarray(aa, dtype=np.float64)
        angle = math.sqrt(float(aa.dot(aa)))
        if angle < 1e-6:
            angle, axis = 0.0, np.array([1.0, 0.0, 0.0])
        else:
            axis = aa / angle
            axis = axis / math.sqrt(float(axis.dot(axis)))
        return np.concatenate([[math.cos(angle / 2)], math.sin(angle / 2) * axis])
    pm, gm = np.asarray(predicted_motion, dtype=np.float64), np.asarray(gt_motion, dtype=np.float64)
    d = abs(float(quat(gm[0:3]).dot(quat(pm[0:3]))))
    rotation_angle_dist = 0.0 if d >= 1.0 else 2.0 * math.acos(d)
    gt_trans, pred_trans = gm[3:6].copy(), pm[3:6].copy()
    if normalize_translations:
        gt_trans = gt_trans / math.sqrt(float(gt_trans.dot(gt_trans)))
        if math.sqrt(float(pred_trans.dot(pred_trans))) > 1e-6:
            pred_trans = pred_trans / math.sqrt(float(pred_trans.dot(pred_trans)))
    diff = gt_trans - pred_trans
    translation_dist = math.sqrt(float(diff.dot(diff)))
    translation_angle_diff = math.acos(float(np.clip(gt_trans.dot(pred_trans), -1, 1)))
    return float(np.rad2deg(rotation_angle_dist)), translation_dist, float(np.rad2deg(translation_angle_diff))
