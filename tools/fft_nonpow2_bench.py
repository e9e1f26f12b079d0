"""Time ffcb_rfft2 + ffcb_irfft2 on planes without a compile-time FFT plan (row f2), direct DFT vs runtime
mixed-radix Stockham, CUDA events on the launch stream.  Run on the GPU box:  python tools/fft_nonpow2_bench.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lama_b200 import _lib as L          # noqa: E402
from lama_b200 import engine as E        # noqa: E402


def time_pair(b, c, h, w, mixed, reps=10):
    os.environ["FFCB_FFT_MIXED_RADIX"] = mixed
    wf = w // 2 + 1
    prog = E.Program("fft_bench", L.MATH_FP32)
    X = prog.buf("x", b, h, w, c); S = prog.buf("s", b, h, wf, 2 * c); O = prog.buf("o", b, h, w, c)
    prog.ops += [E.RfftOp(E.TV(X), E.TV(S)), E.IrfftOp(E.TV(S), E.TV(X), E.TV(O))]
    ex = E.CudaExecutor(prog, torch.device("cuda:0"))
    ex.storage[X.name].normal_()
    for _ in range(3):
        ex.run({})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ex.run({})
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = 4 * b * h * w * c * 3                      # read x, read residual, write out (spectrum not counted)
    return {"plane": f"{h}x{w}", "B": b, "C": c, "mixed_radix": mixed, "ms": round(ms, 4),
            "alg_GBps": round(alg / ms / 1e6, 1), "launches": ex.launches_per_run}


if __name__ == "__main__":
    for (h, w) in [(96, 128), (125, 188), (135, 240), (64, 64)]:
        for mixed in ("0", "1"):
            print(json.dumps(time_pair(8, 192, h, w, mixed)), flush=True)
