"""Runs ONE convolution shape of the DeMoN graphs at batch 64 through the C-ABI test entry (for ncu captures)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demon_b200 import _lib

SHAPES = {
    # name: (kind, B, H, W, Cin, Cout, kh, kw, sy, sx)
    "pd0_conv1": ("conv", 64, 192, 256, 64, 16, 3, 3, 1, 1),
    "refine_conv1_1": ("conv", 64, 96, 128, 64, 64, 3, 3, 1, 1),
    "refine_conv2_1": ("conv", 64, 48, 64, 128, 128, 3, 3, 1, 1),
    "refine0_upconv": ("deconv", 64, 96, 128, 128, 32),
    "conv1y": ("conv", 64, 192, 256, 8, 32, 9, 1, 2, 1),
    "conv1x": ("conv", 64, 96, 256, 32, 32, 1, 9, 1, 2),
    "conv2y": ("conv", 64, 96, 128, 32, 64, 7, 1, 2, 1),
    "predict2_conv1": ("conv", 64, 48, 64, 128, 24, 3, 3, 1, 1),
    "conv5_1y": ("conv", 64, 6, 8, 512, 512, 3, 1, 1, 1),
    "conv3x": ("conv", 64, 24, 64, 128, 128, 1, 5, 1, 2),
    "refine2_upconv": ("deconv", 64, 24, 32, 256, 64),
    "refine3_upconv": ("deconv", 64, 12, 16, 576, 128),   # netFlow2/refine3/upconv: the 514-channel concat4 padded to 18 chunks
    "refine4_upconv": ("deconv", 64, 6, 8, 512, 256),
    "conv2_1y": ("conv", 64, 48, 64, 64, 64, 3, 1, 1, 1),
    "conv4x": ("conv", 64, 12, 32, 256, 256, 1, 5, 1, 2),
    "conv5_1x": ("conv", 64, 6, 8, 512, 512, 1, 3, 1, 1),
}


def main():
    name = sys.argv[1]
    prec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    lib = _lib.load()
    spec = SHAPES[name]
    rng = np.random.RandomState(0)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    timing = os.environ.get("TC_TIMING") == "1"
    if timing:
        lib.demon_debug_tc_timing(1, None, 0)
    if spec[0] == "conv":
        _, B, H, W, Cin, Cout, kh, kw, sy, sx = spec
        x = torch.rand(B, H, W, Cin, device="cuda") - 0.5
        out = torch.empty(B, -(-H // sy), -(-W // sx), Cout, device="cuda")
        k = (rng.standard_normal((kh, kw, Cin, Cout)) / np.sqrt(kh * kw * Cin)).astype(np.float32)
        b = np.zeros(Cout, np.float32)
        for _ in range(reps):
            _lib.check(lib.demon_conv2d_nhwc(x.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, kh, kw, sy, sx, k.ctypes.data, b.ctypes.data, 1, prec, stream))
    else:
        _, B, H, W, Cin, Cout = spec
        x = torch.rand(B, H, W, Cin, device="cuda") - 0.5
        out = torch.empty(B, 2 * H, 2 * W, Cout, device="cuda")
        k = (rng.standard_normal((4, 4, Cout, Cin)) / np.sqrt(4 * Cin)).astype(np.float32)
        b = np.zeros(Cout, np.float32)
        for _ in range(reps):
            _lib.check(lib.demon_deconv4x4s2_nhwc(x.data_ptr(), out.data_ptr(), B, H, W, Cin, Cout, k.ctypes.data, b.ctypes.data, 1, prec, stream))
    torch.cuda.synchronize()
    print(name, "done; timeouts", lib.demon_debug_tc_timeouts())
    print("  kernel time of the last call: %.4f ms" % lib.demon_debug_last_conv_ms())
    if timing:
        full = np.zeros((256, 16), np.int64)
        if lib.demon_debug_tc_timing(0, full.ctypes.data, 256) == 0:
            buf = full[:148]
            names = ["A-producer wait A_empty", "W-producer wait W_empty", "MMA wait accum_empty", "MMA wait T_full", "stager0 wait T_empty",
                     "MMA wait W_full", "stager0 wait A_full", "epilogue wait accum_full", "A-producer total", "MMA total", "stager0 total",
                     "epilogue total", "stager1 wait T_empty", "stager0: address math + shared loads issued", "stager0: tcgen05.st x2 + split + wait::st", "stager0: fence + arrive"]
            m = buf.mean(axis=0)
            for i, nm in enumerate(names):
                print("  %-44s %12.0f cycles (avg per CTA)" % (nm, m[i]))
            ev = full.reshape(-1)[160 * 16:160 * 16 + 640].reshape(10, 64)
            if ev.any():   # diagnostic build: event times of CTA 0, first 64 global steps of the LAST launch
                t0 = ev[:9][ev[:9] > 0].min()
                print("  step | stager: start  free[] seen  wait::st done  arrived | MMA: step entered  rdy  full[] seen  fenced  MMAs issued  commit issued   (cycles since the first event)")
                for sidx in range(24, 48):
                    obs, com, fre, arr, st0, stw, fen, iss, ent = (int(ev[k][sidx] - t0) if ev[k][sidx] else -1 for k in range(9))
                    print("  %4d | %13d %12d %14d %8d | %17d %4d %12d %7d %12d %14d" % (sidx, st0, fre, stw, arr, ent, int(ev[9][sidx]), obs, fen, iss, com))


if __name__ == "__main__":
    main()
