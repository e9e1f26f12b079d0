#!/usr/bin/env python
"""Time the four contractions of one big-lama residual block (bs32, 64x64) under the FFCB_TC_DEBUG knobs
to see which stage of the tcgen05 pipeline bounds each (epilogue / MMA / operand loads)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

os.environ["LAMA_B200_MATH"] = "bf16x3"
from lama_b200 import _lib as L, engine as E, modules as M  # noqa: E402
from lama_b200.testing import BIG_LAMA_KWARGS, seeded_parameters_, synthetic_image_mask, generator_input  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
gen = seeded_parameters_(M.FFCResNetGenerator(**BIG_LAMA_KWARGS).eval(), 0).to(dev)
img, mask = synthetic_image_mask(B, 512, 0)
x = generator_input(img, mask).to(dev)
ex = E.get_executor(gen, "generator", (x,), math=L.MATH_BF16X3)
ex.run({"x0": x})
torch.cuda.synchronize()
names = [n for n, _f, _a in ex.calls]
# ops of the 9th residual block (second FFC_BN_ACT: with residual addends)
want = ["ffcb_conv:convl2l+convg2l+bn_l+act", "ffcb_conv:st.conv1+bn+relu", "ffcb_conv:fu.conv_layer+bn+relu",
        "ffcb_conv:convl2g+st.conv2+bn_g+act", "ffcb_rfft2", "ffcb_irfft2"]
if os.environ.get("TC_OPS"):       # e.g. TC_OPS="stem 7x7,head 7x7 rows,convT phase 11": last call whose name contains each
    want = []
    for sub in os.environ["TC_OPS"].split(","):
        cand = [n for n in names if sub in n]
        assert cand, (sub, sorted(set(names)))
        want.append(cand[-1])
idx = {}
for w in want:
    cand = [i for i, n in enumerate(names) if n == w]
    idx[w] = cand[len(cand) // 2 + 1] if len(cand) > 2 else cand[-1]
stream = torch.cuda.current_stream(dev).cuda_stream


def time_call(i, reps=20):
    n, fn, a = ex.calls[i]
    for _ in range(3):
        fn(*a, stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        rc = fn(*a, stream)
        assert rc == 0, L.get_lib().ffcb_last_error()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


if os.environ.get("ONCE"):
    # one launch of each op inside a cudaProfilerStart/Stop range: `ncu --profile-from-start off --set full ...`
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for w in want:
        n, fn, a = ex.calls[idx[w]]
        assert fn(*a, stream) == 0
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("once:", [names[idx[w]] for w in want])
    sys.exit(0)
if os.environ.get("SWEEP_BN"):
    print("N-tile sweep (FFCB_TC_BN), us per launch")
    for w in want[:4]:
        row = []
        for bn in ("32", "64", "96", "128"):
            os.environ["FFCB_TC_BN"] = bn
            row.append(time_call(idx[w]))
        del os.environ["FFCB_TC_BN"]
        row.append(time_call(idx[w]))
        print(f"{w:45s} " + " ".join(f"{v:10.1f}" for v in row) + "   (32 64 96 128 default)")
    sys.exit(0)
print(f"{'op':45s} " + " ".join(f"{k:>10s}" for k in ["full", "noGlobal", "noEpi", "noMMA", "noA", "noMMA+Epi", "noA+noEpi"]))
for w in want:
    row = []
    for dbg in ([0, 1, 2, 4, 8, 6, 10] if w.startswith("ffcb_conv") else [0]):
        os.environ["FFCB_TC_DEBUG"] = str(dbg)
        row.append(time_call(idx[w]))
    os.environ["FFCB_TC_DEBUG"] = "0"
    if w == "ffcb_irfft2":
        for variant in ("1", "2"):
            os.environ["FFCB_FFT_INV_PLANE"] = variant
            row.append(time_call(idx[w]))
        del os.environ["FFCB_FFT_INV_PLANE"]
    if w in ("ffcb_rfft2", "ffcb_irfft2"):
        os.environ["FFCB_FFT_TWO_PASS"] = "1"
        row.append(time_call(idx[w]))
        del os.environ["FFCB_FFT_TWO_PASS"]
    print(f"{w:45s} " + " ".join(f"{v:10.1f}" for v in row))
