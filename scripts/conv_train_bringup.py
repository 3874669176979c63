"""Bring-up of the training convolutions (csrc/conv_train.cu) on a GPU box: fprop with the MN-major activation operand
wgrad, moments; errors against float64.
    python scripts/conv_train_bringup.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import conv_train as ct  # noqa: E402


def rel(a, b):
    return ((a.double().cpu() - b.cpu()).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def fprop_case(B, K, L, rows, passes):
    g = torch.Generator().manual_seed(B * 1000 + K + L + rows)
    x = torch.randn(B, K, L, generator=g)
    w = torch.randn(rows, K, generator=g) / K ** 0.5
    want = torch.einsum("rk,bkl->brl", w.double(), x.double())
    hi, lo = ct.split_planes(x.cuda())
    a_hi, a_lo = ct.split_weight(w.cuda())
    out, mom = ct.conv1x1(hi, lo, a_hi, a_lo, rows, K, want_moments=True, passes=passes)
    torch.cuda.synchronize()
    e = rel(out, want)
    s1 = want.sum(dim=(0, 2)); s2 = (want * want).sum(dim=(0, 2))
    em = max(rel(mom[:, 0], s1), rel(mom[:, 1], s2))
    return e, em


def wgrad_case(B, Co, Ci, L, passes):
    g = torch.Generator().manual_seed(B + Co + Ci + L)
    gz = torch.randn(B, Co, L, generator=g)
    x = torch.randn(B, Ci, L, generator=g)
    want = torch.einsum("bol,bil->oi", gz.double(), x.double())
    g_hi, g_lo = ct.split_planes(gz.cuda())
    x_hi, x_lo = ct.split_planes(x.cuda())
    dw = ct.wgrad(g_hi, g_lo, x_hi, x_lo, passes=passes)
    torch.cuda.synchronize()
    return rel(dw, want)


def main():
    torch.cuda.set_device(0)
    for swap in ("0",):
        for (B, K, L, rows) in [(1, 64, 256, 128), (2, 128, 512, 128), (2, 6, 320, 128), (3, 259, 1280, 256), (1, 128, 1024, 1),
                                (2, 1536, 256, 1024)]:
            try:
                e3, em = fprop_case(B, K, L, rows, 3)
                e1, _ = fprop_case(B, K, L, rows, 1)
                print(f"swap={swap} fprop B={B} K={K} L={L} rows={rows}: rel err passes3 {e3:.3e} passes1 {e1:.3e} moments {em:.3e}", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"swap={swap} fprop B={B} K={K} L={L} rows={rows}: FAILED {ex}", flush=True)
                return
    for (B, Co, Ci, L) in [(1, 128, 128, 256), (2, 128, 6, 320), (3, 256, 259, 1280), (2, 1024, 1536, 256), (2, 1, 128, 1024),
                           (15, 128, 128, 20480)]:
        try:
            e3 = wgrad_case(B, Co, Ci, L, 3)
            e1 = wgrad_case(B, Co, Ci, L, 1)
            print(f"wgrad B={B} Co={Co} Ci={Ci} L={L}: rel err passes3 {e3:.3e} passes1 {e1:.3e}", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"wgrad B={B} Co={Co} Ci={Ci} L={L}: FAILED {ex}", flush=True)
            return
    # timing at the SA level-0 shape (15 x 128 x 327680)
    B, K, L, rows = 15, 128, 327680, 128
    x = torch.randn(B, K, L, device="cuda")
    w = torch.randn(rows, K, device="cuda") / K ** 0.5
    hi, lo = ct.split_planes(x)
    a_hi, a_lo = ct.split_weight(w)
    del x
    for passes in (3, 1):
        for _ in range(2):
            out = ct.conv1x1(hi, lo, a_hi, a_lo, rows, K, want_moments=True, passes=passes)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = ct.conv1x1(hi, lo, a_hi, a_lo, rows, K, want_moments=True, passes=passes)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = (B * K * L * 4 + B * rows * L * 4) / 1e9
        print(f"fprop 15x128x327680 -> 128, passes {passes}: {ms:.3f} ms, {gb / ms:.1f} TB/s... ({gb / (ms * 1e-3):.0f} GB/s), "
              f"{2 * B * K * L * rows * passes / ms / 1e9:.0f} TFLOP/s executed", flush=True)
    g_hi, g_lo = ct.split_planes(out[0] if isinstance(out, tuple) else out)
    for passes in (3, 1):
        for _ in range(2):
            dw = ct.wgrad(g_hi, g_lo, hi, lo, passes=passes)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dw = ct.wgrad(g_hi, g_lo, hi, lo, passes=passes)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = (B * K * L * 4 + B * rows * L * 4) / 1e9
        print(f"wgrad 15x128x327680, passes {passes}: {ms:.3f} ms ({gb / (ms * 1e-3):.0f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
