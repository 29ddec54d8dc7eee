"""Generates tests/golden/metrics_golden.npz by running the REFERENCE's own metric functions
(/root/reference/python/depthmotionnet/evaluation/metrics.py, imported unmodified) on seeded inputs.  The module imports
`minieigen` (absent here) for compute_motion_errors only; a minimal stand-in (Vector3, Quaternion(angle, axis),
angularDistance = 2*acos(min(1,|q1.q2|)) as in Eigen) is injected before the import.  Run in the build container:
    python tests/golden/make_metrics_golden.py
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np

REF = "/root/reference/python/depthmotionnet/evaluation/metrics.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_golden.npz")


class Vector3:
    def __init__(self, x, y, z):
        self.v = np.array([x, y, z], dtype=np.float64)

    def norm(self):
        return math.sqrt(float(self.v.dot(self.v)))

    def normalize(self):
        self.v = self.v / self.norm()

    def dot(self, o):
        return float(self.v.dot(o.v))

    def __sub__(self, o):
        return Vector3(*(self.v - o.v))


class Quaternion:
    def __init__(self, angle, axis):
        s = math.sin(angle / 2)
        self.q = np.array([math.cos(angle / 2), s * axis.v[0], s * axis.v[1], s * axis.v[2]])

    def angularDistance(self, o):
        d = abs(float(self.q.dot(o.q)))
        return 0.0 if d >= 1.0 else 2.0 * math.acos(d)


def main():
    stub = types.ModuleType("minieigen")
    stub.Vector3, stub.Quaternion = Vector3, Quaternion
    sys.modules["minieigen"] = stub
    spec = importlib.util.spec_from_file_location("ref_metrics", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)

    rng = np.random.RandomState(2024)
    out = {}
    names = ['l1', 'l1_inverse', 'scale_invariant', 'abs_relative', 'sq_relative', 'avg_log10', 'rmse_log', 'rmse',
             'ratio_threshold_1.25', 'ratio_threshold_1.5625', 'ratio_threshold_1.953125']
    cases = []
    for ci, (h, w) in enumerate([(48, 64), (192, 256), (7, 5)]):
        gt = rng.uniform(0.05, 2.0, (h, w)).astype(np.float32)                       # inverse depths
        pred = (gt * rng.uniform(0.6, 1.6, (h, w)) * 0.8).astype(np.float32)
        # invalid pixels of every kind (metrics.py:25-38)
        bad = rng.rand(h, w)
        gt[bad < 0.03] = np.nan; gt[(bad >= 0.03) & (bad < 0.05)] = 0.0; gt[(bad >= 0.05) & (bad < 0.06)] = -1.0
        pred[(bad >= 0.06) & (bad < 0.08)] = np.inf; pred[(bad >= 0.08) & (bad < 0.09)] = -0.5
        t = np.array([0.9, 0.1, -0.05]) * (1.0 if ci == 1 else 1.7)
        if ci == 1:
            t = t / np.sqrt(t.dot(t))                                                  # normalised translation: no gt scaling
        for scaling in ('abs', 'log', 'inv'):
            errs, errs_scaled = m.evaluate_depth(t, gt, pred, depth_scaling=scaling)
            cases.append((ci, scaling, errs, errs_scaled))
        out["gt_%d" % ci], out["pred_%d" % ci], out["t_%d" % ci] = gt, pred, t
        # plain compute_errors on depths (not inverse)
        d_gt, d_pred = rng.uniform(0.5, 10, (h, w)).astype(np.float32), None
        d_pred = (d_gt * rng.uniform(0.7, 1.4, (h, w))).astype(np.float32)
        d_pred[bad < 0.04] = np.nan
        e = m.compute_errors(d_pred, d_gt)
        out["dgt_%d" % ci], out["dpred_%d" % ci] = d_gt, d_pred
        out["errs_plain_%d" % ci] = np.array([e['num_valid']] + [e[k] for k in names], dtype=np.float64)
        f1 = rng.uniform(-0.2, 0.2, (2, h, w)).astype(np.float32)
        f2 = (f1 + rng.normal(0, 0.01, (2, h, w))).astype(np.float32)
        f2[0, bad < 0.05] = np.nan
        f2[:, (bad > 0.5) & (bad < 0.52)] = f1[:, (bad > 0.5) & (bad < 0.52)]       # epe == 0 is masked out (valid mask needs > 0)
        out["f1_%d" % ci], out["f2_%d" % ci] = f1, f2
        out["epe_%d" % ci] = np.float64(m.compute_flow_epe(f1, f2))
    for ci, scaling, errs, errs_scaled in cases:
        out["errs_%d_%s" % (ci, scaling)] = np.array([errs['num_valid']] + [errs[k] for k in names], dtype=np.float64)
        out["errs_scaled_%d_%s" % (ci, scaling)] = np.array([errs_scaled['num_valid']] + [errs_scaled[k] for k in names], dtype=np.float64)
    # all-invalid input
    e = m.compute_errors(np.full((4, 4), np.nan, np.float32), np.ones((4, 4), np.float32))
    out["errs_all_invalid"] = np.array([e['num_valid']] + [e[k] for k in names], dtype=np.float64)
    motions = rng.uniform(-0.3, 0.3, (6, 6))
    motions[:, 3:] += np.array([0.9, 0.1, -0.05])
    motions[2, :3] = 1e-9
    res = []
    for i in range(0, 6, 2):
        for nt in (True, False):
            res.append(m.compute_motion_errors(motions[i], motions[i + 1], nt))
    out["motions"], out["motion_errors"] = motions, np.array(res, dtype=np.float64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
