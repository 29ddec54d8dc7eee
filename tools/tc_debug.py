"""GPU debugging aid for the tcgen05 convolution kernel: runs a ladder of cases from trivial to full
through the C-ABI test entry and prints error structure (which rows / channels / taps are wrong)."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demon_b200 import _lib

lib = _lib.load()


def run(x, k, b, sy, sx, leaky, prec):
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = k.shape
    xin = torch.from_numpy(x).cuda()
    Ho, Wo = -(-H // sy), -(-W // sx)
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda")
    kk, bb = np.ascontiguousarray(k, np.float32), np.ascontiguousarray(b, np.float32)
    rc = lib.demon_conv2d_nhwc(xin.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, kh, kw, sy, sx, kk.ctypes.data, bb.ctypes.data,
                               int(leaky), prec, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        return None, lib.demon_last_error().decode()
    torch.cuda.synchronize()
    return out.cpu().numpy(), None


def ref(x, k, b, sy, sx, leaky):
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2)
    kt = torch.from_numpy(k).double().permute(3, 2, 0, 1)
    kh, kw = k.shape[:2]
    y = F.conv2d(F.pad(xt, (kw // 2, kw // 2, kh // 2, kh // 2)), kt, torch.from_numpy(b).double(), stride=(sy, sx))
    if leaky:
        y = torch.maximum(0.1 * y, y)
    return y.permute(0, 2, 3, 1).numpy()


def run_deconv(x, k, b, leaky, prec):
    B, H, W, Cin = x.shape
    Cout = k.shape[2]
    xin = torch.from_numpy(x).cuda()
    out = torch.full((B, 2 * H, 2 * W, Cout), float("nan"), device="cuda")
    kk, bb = np.ascontiguousarray(k, np.float32), np.ascontiguousarray(b, np.float32)
    rc = lib.demon_deconv4x4s2_nhwc(xin.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, kk.ctypes.data, bb.ctypes.data, int(leaky), prec,
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        return None, lib.demon_last_error().decode()
    torch.cuda.synchronize()
    return out.cpu().numpy(), None


def ref_deconv(x, k, b, leaky):
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2)
    kt = torch.from_numpy(k).double().permute(3, 2, 0, 1)
    y = F.conv_transpose2d(xt, kt, torch.from_numpy(b).double(), stride=2, padding=1)
    if leaky:
        y = torch.maximum(0.1 * y, y)
    return y.permute(0, 2, 3, 1).numpy()


def report(name, got, want):
    if got is None:
        print("%-44s ERROR" % name)
        return
    nan = np.isnan(got)
    err = np.abs(np.where(nan, 0, got) - want)
    rel = err.max() / np.abs(want).max()
    print("%-44s rel_err %.3e  nan %d/%d  timeouts %d" % (name, rel, nan.sum(), got.size, lib.demon_debug_tc_timeouts()))
    if rel > 1e-2 or nan.any():
        bad = (err > 1e-2 * np.abs(want).max()) | nan
        B, H, W, C = got.shape
        print("    bad fraction %.3f; by channel%%8: %s" % (bad.mean(), np.round([bad[..., c::8].mean() for c in range(min(8, C))], 2)))
        print("    by x%%8: %s" % np.round([bad[:, :, xx::8].mean() for xx in range(min(8, W))], 2))
        print("    by y: %s" % np.round([bad[:, yy].mean() for yy in range(min(H, 12))], 2))
        print("    by n: %s" % np.round([bad[nn].mean() for nn in range(B)], 2))
        print("    got[0,0,0,:4] %s want %s" % (got[0, 0, 0, :4], want[0, 0, 0, :4]))


def main():
    rng = np.random.RandomState(0)
    halo_cases = [
        ("CIN8 9x1 s2 32x16->16x16 Cout32", 2, 32, 16, 8, 32, 9, 1, 2, 1),
        ("CIN8 3x3 16x16 Cout32", 1, 16, 16, 8, 32, 3, 3, 1, 1),
        ("CIN8 9x1 s2 192x256 Cout32 B2", 2, 192, 256, 8, 32, 9, 1, 2, 1),
        ("HALO 1x1 16x8 Cin32 Cout32", 1, 16, 8, 32, 32, 1, 1, 1, 1),
        ("HALO 3x1 16x8 Cin32 Cout32", 1, 16, 8, 32, 32, 3, 1, 1, 1),
        ("HALO 1x3 16x8 Cin32 Cout32", 1, 16, 8, 32, 32, 1, 3, 1, 1),
        ("HALO 3x3 32x16 Cin64 Cout64", 2, 32, 16, 64, 64, 3, 3, 1, 1),
        ("HALO 3x3 48x64 Cin128 Cout24", 2, 48, 64, 128, 24, 3, 3, 1, 1),
        ("HALO 3x3 32x24 Cin64 Cout16", 1, 32, 24, 64, 16, 3, 3, 1, 1),
        ("HALO 1x9 s2 32x32->32x16 Cin32 Cout32", 2, 32, 32, 32, 32, 1, 9, 1, 2),
        ("HALO 7x1 s2 64x16->32x16 Cin32 Cout32", 2, 64, 16, 32, 32, 7, 1, 2, 1),
        ("HALO 1x7 s2 Cin32 Cout32", 1, 16, 32, 32, 32, 1, 7, 1, 2),
        ("HALO 3x3 Cin64 Cout128 48x64", 1, 48, 64, 64, 128, 3, 3, 1, 1),
    ]
    for name, B, H, W, Cin, Cout, kh, kw, sy, sx in halo_cases:
        x = rng.uniform(-1, 1, (B, H, W, Cin)).astype(np.float32)
        k = (rng.standard_normal((kh, kw, Cin, Cout)) / np.sqrt(kh * kw * Cin)).astype(np.float32)
        b = rng.uniform(-0.1, 0.1, Cout).astype(np.float32)
        want = ref(x, k, b, sy, sx, True)
        for prec, pname in ((2, "tf32"), (1, "3xtf32")):
            got, e = run(x, k, b, sy, sx, True, prec)
            if e:
                print("%-44s %s" % (name + " " + pname, e))
                continue
            report(name + " " + pname, got, want)
        if lib.demon_debug_tc_timeouts():
            print("pipeline timeout flagged -- stopping")
            return
    for name, B, H, W, Cin, Cout in [("HALO deconv 16x8 Cin32 Cout32", 1, 16, 8, 32, 32), ("HALO deconv 48x64 Cin128 Cout64", 2, 48, 64, 128, 64),
                                     ("HALO deconv 32x32 Cin128 Cout32", 1, 32, 32, 128, 32), ("per-tap deconv 24x32 Cin256 Cout64", 2, 24, 32, 256, 64),
                                     ("per-tap deconv 12x16 Cin544 Cout128", 1, 12, 16, 544, 128), ("per-tap deconv 6x8 Cin512 Cout256", 2, 6, 8, 512, 256)]:
        x = rng.uniform(-1, 1, (B, H, W, Cin)).astype(np.float32)
        k = (rng.standard_normal((4, 4, Cout, Cin)) / np.sqrt(4 * Cin)).astype(np.float32)
        b = rng.uniform(-0.1, 0.1, Cout).astype(np.float32)
        want = ref_deconv(x, k, b, True)
        for prec, pname in ((2, "tf32"), (1, "3xtf32")):
            got, e = run_deconv(x, k, b, True, prec)
            if e:
                print("%-44s %s" % (name + " " + pname, e))
                continue
            report(name + " " + pname, got, want)
        if lib.demon_debug_tc_timeouts():
            print("pipeline timeout flagged -- stopping")
            return
    cases = [
        # name, B,H,W,Cin,Cout,kh,kw,sy,sx
        ("1x1 Cin32 Cout32 128px", 1, 2, 64, 32, 32, 1, 1, 1, 1),
        ("1x1 Cin32 Cout16", 1, 2, 64, 32, 16, 1, 1, 1, 1),
        ("1x1 Cin64 Cout64 (2 chunks)", 1, 2, 64, 64, 64, 1, 1, 1, 1),
        ("1x1 Cin32 Cout128 2 tiles M", 1, 4, 64, 32, 128, 1, 1, 1, 1),
        ("1x1 Cin32 Cout512 (2 n-tiles)", 1, 2, 64, 32, 512, 1, 1, 1, 1),
        ("3x1 Cin32 Cout32", 1, 4, 64, 32, 32, 3, 1, 1, 1),
        ("1x3 Cin32 Cout32", 1, 4, 64, 32, 32, 1, 3, 1, 1),
        ("3x3 Cin64 Cout64", 2, 8, 16, 64, 64, 3, 3, 1, 1),
        ("5x1 s2 Cin64 Cout128", 2, 24, 32, 64, 128, 5, 1, 2, 1),
        ("1x5 s2 Cin128 Cout256", 3, 12, 16, 128, 256, 1, 5, 1, 2),
        ("3x3 s2 Cin32 Cout64", 1, 16, 24, 32, 64, 3, 3, 2, 2),
        ("3x1 6x8 images Cin512 Cout512", 3, 6, 8, 512, 512, 3, 1, 1, 1),
        ("3x3 odd size 19x21", 1, 19, 21, 64, 64, 3, 3, 1, 1),
        ("many tiles persistent 3x3 Cin64 Cout32", 4, 96, 128, 64, 32, 3, 3, 1, 1),
    ]
    for name, B, H, W, Cin, Cout, kh, kw, sy, sx in cases:
        x = rng.uniform(-1, 1, (B, H, W, Cin)).astype(np.float32)
        k = (rng.standard_normal((kh, kw, Cin, Cout)) / np.sqrt(kh * kw * Cin)).astype(np.float32)
        b = rng.uniform(-0.1, 0.1, Cout).astype(np.float32)
        want = ref(x, k, b, sy, sx, True)
        for prec, pname in ((2, "tf32"), (1, "3xtf32")):
            got, e = run(x, k, b, sy, sx, True, prec)
            if e:
                print("%-44s %s" % (name + " " + pname, e))
                continue
            report(name + " " + pname, got, want)
        got, _ = run(x, k, b, sy, sx, True, 0)
        report(name + " simt", got, want)
        if lib.demon_debug_tc_timeouts():
            print("pipeline timeout flagged -- stopping")
            break


if __name__ == "__main__":
    main()
