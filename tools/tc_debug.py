#!/usr/bin/env python
"""Bring-up harness for the tcgen05 arm (conv_tc.cu): runs graded ffcb_conv cases in FFCB_MATH_BF16X3,
each in its own subprocess with a timeout (a deadlocked pipeline must not take the others down), and
prints error structure (by accumulator row, by output channel, by K block) for offline diagnosis.

    python tools/tc_debug.py            # all cases
    python tools/tc_debug.py --case N   # one case in-process
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, B, H, W, Cin, N, kind
    ("flat_1x1_single_tile", 1, 8, 16, 64, 64, "flat"),
    ("flat_1x1_k128_n192", 1, 8, 16, 128, 192, "flat"),
    ("flat_1x1_w33_multi_tile", 2, 16, 33, 128, 384, "flat"),
    ("k3_reflect_16x16", 2, 16, 16, 64, 128, "k3"),
    ("k3_reflect_64x64_multiwave", 6, 64, 64, 128, 384, "k3"),
    ("k3_stride2", 1, 32, 32, 64, 64, "k3s2"),
    ("zero_border_phases", 2, 8, 8, 64, 64, "convT"),
    ("two_sources_addend_post", 2, 8, 8, 128, 64, "two"),
    ("ragged_channels_pad", 2, 16, 16, 24, 40, "k3"),
    ("resblock_shape_L", 2, 64, 64, 512, 128, "k3"),
]


def run_case(idx):
    import numpy as np
    import torch
    from lama_b200 import _lib as L
    from lama_b200 import engine as E
    from lama_b200 import packing as P

    name, b, h, w, cin, n, kind = CASES[idx]
    g = torch.Generator().manual_seed(100 + idx)
    rn = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    q = lambda t: (lambda s: s[0].float() + s[1].float())(P.split_bf16(t))   # value the kernel actually sees
    prog = E.Program("tc_debug", L.MATH_BF16X3)
    feed = {}

    def add_input(nm, t, halo):
        bb = prog.buf(nm, *t.shape, gemm=True, halo=halo)
        prog.inputs[nm] = (t.shape[0], t.shape[3], t.shape[1], t.shape[2])
        prog.ops.append(E.ToNHWC(nm, E.TV(bb)))
        feed[nm] = t.permute(0, 3, 1, 2).contiguous()
        return bb

    post, add, out_hw = False, None, (h, w)
    x = q(rn(b, h, w, cin))
    ins_ref = [x, None]
    if kind == "flat":
        pk = P.pack_conv([(rn(n, cin, 1, 1) * 0.1, 0, 0, 0)], None, rn(n), act=L.ACT_NONE)
        X = add_input("x0", x, halo=False); tvs = [E.TV(X), None]
    elif kind == "k3":
        pk = P.pack_conv([(rn(n, cin, 3, 3) * 0.1, 0, 0, 1)], rn(n).abs() + 0.5, rn(n), act=L.ACT_RELU)
        X = add_input("x0", x, halo=True); tvs = [E.TV(X), None]
    elif kind == "k3s2":
        pk = P.pack_conv([(rn(n, cin, 3, 3) * 0.1, 0, 0, 1)], None, rn(n), stride=2, act=L.ACT_RELU)
        X = add_input("x0", x, halo=True); tvs = [E.TV(X), None]; out_hw = (h // 2, w // 2)
    elif kind == "two":
        x1 = q(rn(b, h, w, 192)); ins_ref = [x, x1]
        pk = P.pack_conv([(rn(n, 64, 3, 3) * 0.1, 0, 64, 1), (rn(n, 192, 1, 1) * 0.1, 1, 0, 0)], rn(n).abs(), rn(n),
                         act=L.ACT_RELU)
        X = add_input("x0", x, halo=True); X1 = add_input("x1", x1, halo=False); tvs = [E.TV(X), E.TV(X1)]
        add, post = rn(b, h, w, n), True
    if kind == "convT":
        wt = rn(cin, n, 3, 3) * 0.1
        phases = P.pack_conv_transpose_phases(wt, rn(n), rn(n).abs() + 0.5, rn(n), act=L.ACT_RELU)
        X = add_input("x0", x, halo=True)
        Y = prog.buf("y", b, 2 * h, 2 * w, n)
        want = torch.zeros(b, 2 * h, 2 * w, n, dtype=torch.float64)
        for a, bb, pk in phases:
            pk.w_kn = q(pk.w_kn)
            prog.ops.append(E.ConvOp(pk, [E.TV(X), None], E.TV(Y, phase=(a, bb))))
            want[:, a::2, bb::2] = P.apply_packed_reference(pk, [x, None], (h, w))
        out_shape = (b, n, 2 * h, 2 * w)
    else:
        pk.w_kn = q(pk.w_kn)      # make the weights exactly representable too: isolates kernel bugs from rounding
        Y = prog.buf("y", b, out_hw[0], out_hw[1], n)
        atv = None
        if add is not None:
            A = prog.buf("add", *add.shape)
            prog.inputs["xa"] = (b, n, out_hw[0], out_hw[1])
            prog.ops.append(E.ToNHWC("xa", E.TV(A))); feed["xa"] = add.permute(0, 3, 1, 2).contiguous(); atv = E.TV(A)
        prog.ops.append(E.ConvOp(pk, tvs, E.TV(Y), addend=atv, addend_post=post))
        want = P.apply_packed_reference(pk, ins_ref, out_hw, addend=add, addend_post=post)
        out_shape = (b, n, out_hw[0], out_hw[1])
    prog.ops.append(E.ToNCHW(E.TV(Y), "y0")); prog.outputs = {"y0": out_shape}
    E.insert_border_ops(prog)
    ex = E.CudaExecutor(prog, torch.device("cuda:0"))
    out = ex.run({k: v.cuda() for k, v in feed.items()})
    torch.cuda.synchronize()
    got = out["y0"].cpu().permute(0, 2, 3, 1).double()
    err = (got - want).abs()
    scale = float(want.abs().max())
    rel = float(err.max()) / scale
    rep = {"case": name, "rel_err": rel, "scale": scale, "ok": rel < 1e-4,
           "nan": int(torch.isnan(got).sum()), "frac_bad": float((err > 1e-3 * scale).double().mean())}
    if not rep["ok"]:
        bad = err > 1e-3 * scale
        flat = bad.reshape(-1, bad.shape[-1])
        rep["bad_by_channel_first32"] = flat.double().mean(dim=0)[:32].tolist()
        rows = bad.reshape(-1, bad.shape[-1]).any(dim=1).nonzero().flatten()[:40].tolist()
        rep["first_bad_pixels(flat index)"] = rows
        rep["bad_by_pixel_mod128"] = [float(flat.any(dim=1).double()[i::128].mean()) for i in range(0, 128, 8)]
        i = int(err.reshape(-1).argmax())
        rep["worst"] = {"index": list(np.unravel_index(i, tuple(err.shape))), "got": float(got.reshape(-1)[i]),
                        "want": float(want.reshape(-1)[i])}
        rep["sample_got"] = got[0, 0, 0, :8].tolist()
        rep["sample_want"] = want[0, 0, 0, :8].tolist()
        rep["sample_got_p1"] = got[0, 0, 1, :8].tolist()
        rep["sample_want_p1"] = want[0, 0, 1, :8].tolist()
    print("RESULT " + json.dumps(rep))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=None)
    ap.add_argument("--timeout", type=int, default=120)
    args = ap.parse_args()
    if args.case is not None:
        run_case(args.case)
        sys.exit(0)
    for i, c in enumerate(CASES):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i)], capture_output=True,
                               text=True, timeout=args.timeout)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
            if lines:
                print(lines[-1])
            else:
                print(f"RESULT {{\"case\": \"{c[0]}\", \"ok\": false, \"rc\": {r.returncode}, \"stderr\": "
                      f"{json.dumps(r.stderr[-1500:])}}}")
        except subprocess.TimeoutExpired:
            print(f"RESULT {{\"case\": \"{c[0]}\", \"ok\": false, \"timeout\": true}}")
        sys.stdout.flush()
