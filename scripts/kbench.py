"""Kernel micro-benchmark at the BASELINE shapes (one launch per case through depgan_op_conv2d / depgan_op_wgrad,
CUDA events on the launching stream).  Used with ncu (-k regex:...) to read stall reasons and traffic per kernel.

  python scripts/kbench.py [case-substring ...]
"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from depgan_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def act(shape, bf16, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = torch.randn(shape, generator=g).to(dev)
    return t.to(torch.bfloat16) if bf16 else t


def conv_case(name, N, H, W, c0, cout, ks, *, c1=0, tc=True, in_bf16=None, film=False, add=False, mask=False,
              deconv=False, pre=False, relu=True, out_bf16=True, head=0, film_self=False, pool=False):
    in_bf16 = tc if in_bf16 is None else in_bf16
    keep = []
    d = _lib.ConvDesc()
    x = act((N, H, W, c0), in_bf16, 1); keep.append(x)
    d.in0, d.C0, d.C1 = x.data_ptr(), c0, c1
    if c1:
        x1 = act((N, H, W, c1), in_bf16, 2); keep.append(x1)
        d.in1 = x1.data_ptr()
    cin = c0 + c1
    taps = 1 if deconv else ks * ks
    ncols = 4 * cout if deconv else cout
    w = (torch.randn(taps, cin, ncols) / (taps * cin) ** 0.5).to(dev); keep.append(w)
    if tc:
        wb = torch.empty(w.numel(), dtype=torch.bfloat16, device=dev); keep.append(wb)
        if deconv:
            wb.copy_(w.permute(0, 2, 1).reshape(-1).to(torch.bfloat16))
        else:
            _lib.check(L.depgan_op_pack_weights(w.data_ptr(), wb.data_ptr(), taps, cin, cout, st), "pack")
        d.w_bf16 = wb.data_ptr()
    else:
        d.w_f32 = w.data_ptr()
    sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev); keep += [sc, sh]
    d.scale, d.shift = sc.data_ptr(), sh.data_ptr()
    oh, ow = (2 * H, 2 * W) if deconv else (H, W)
    odt = torch.bfloat16 if out_bf16 else torch.float32
    out = torch.empty((N, oh, ow, cout), dtype=odt, device=dev); keep.append(out)
    d.out = out.data_ptr()
    nbytes = x.numel() * x.element_size() * (1 + c1 / max(c0, 1)) + out.numel() * out.element_size()
    if pre:
        p = torch.empty_like(out); keep.append(p); d.out_pre = p.data_ptr(); nbytes += out.numel() * out.element_size()
    if film:
        g, b = torch.ones(N, cout, device=dev), torch.zeros(N, cout, device=dev); keep += [g, b]
        r = x if film_self else act((N, H, W, cout), out_bf16, 3)  # film_self: the residual is the conv input (as in the nets)
        keep.append(r)
        d.film_g, d.film_b, d.film_stride, d.res = g.data_ptr(), b.data_ptr(), cout, r.data_ptr()
        nbytes += r.numel() * r.element_size()
    if add:
        t = act((N, H, W, cout), out_bf16, 4); keep.append(t); d.add_src = t.data_ptr()
        nbytes += t.numel() * t.element_size()
    if mask:
        t = act((N, H, W, cout), out_bf16, 5); keep.append(t); d.mask_src = t.data_ptr()
        nbytes += t.numel() * t.element_size()
    if pool:
        t = torch.empty((N, oh // 2, ow // 2, cout), dtype=odt, device=dev); keep.append(t); d.pool_out = t.data_ptr()
        nbytes += t.numel() * t.element_size()
    if head:
        hw, hb = torch.randn(cout, head, device=dev) * 0.1, torch.zeros(head, device=dev)
        ho = torch.empty((N, H, W, head), device=dev); keep += [hw, hb, ho]
        d.head_w, d.head_b, d.head_out, d.head_nc, d.head_act = hw.data_ptr(), hb.data_ptr(), ho.data_ptr(), head, 1
        nbytes += ho.numel() * 4
    if cout == 1:
        d.scale = d.shift = None
        relu = False
    d.relu, d.deconv = int(relu), int(deconv)
    d.N, d.H, d.W, d.Cout, d.ks = N, H, W, cout, ks
    d.in_bf16, d.out_bf16, d.use_tc = int(in_bf16), int(out_bf16), int(tc)
    flops = 2.0 * N * oh * ow * cout * cin * (1 if deconv else ks * ks)

    def run():
        _lib.check(L.depgan_op_conv2d(C.byref(d), st), name)
    return name, run, flops, nbytes, keep


def wgrad_case(name, N, H, W, cin, cout, ks, mode):
    xb = mode == 1
    x = act((N, H, W, cin), xb, 1)
    dy = act((N, H, W, cout), mode != 0, 2)
    dw = torch.zeros(ks * ks * cin * cout, device=dev)
    flops = 2.0 * N * H * W * cout * cin * ks * ks
    nbytes = x.numel() * x.element_size() + dy.numel() * dy.element_size()

    def run():
        _lib.check(L.depgan_op_wgrad(x.data_ptr(), None, cin, 0, dy.data_ptr(), dw.data_ptr(), N, H, W, cout, ks, mode,
                                     st), name)
    return name, run, flops, nbytes, [x, dy, dw]


CASES = [
    lambda: conv_case("first_3x3_1to32_N64", 64, 256, 256, 1, 32, 3, tc=False, in_bf16=False),
    lambda: conv_case("first_5x5_1to16_N96", 96, 256, 256, 1, 16, 5, tc=False, in_bf16=False),
    lambda: conv_case("first_5x5_1to16_mask_N32", 32, 256, 256, 1, 16, 5, tc=False, in_bf16=False, mask=True),
    lambda: conv_case("last_5x5_16to1_N32", 32, 256, 256, 16, 1, 5, tc=False, in_bf16=True, out_bf16=False),
    lambda: wgrad_case("wgrad_first_5x5_1to16_N64", 64, 256, 256, 1, 16, 5, 2),
    lambda: wgrad_case("wgrad_first_3x3_1to32_N32", 32, 256, 256, 1, 32, 3, 2),
    lambda: conv_case("tc_3x3_32to32_plain_N64", 64, 256, 256, 32, 32, 3),
    lambda: conv_case("tc_3x3_32to32_pool_N64", 64, 256, 256, 32, 32, 3, pool=True),
    lambda: conv_case("tc_3x3_64to64_pool_N64", 64, 128, 128, 64, 64, 3, pool=True),
    lambda: conv_case("tc_3x3_32to32_plain_N8_l2", 8, 256, 256, 32, 32, 3),
    lambda: conv_case("tc_3x3_32to32_plain_N16_l2", 16, 256, 256, 32, 32, 3),
    lambda: conv_case("tc_3x3_32to32_filmA_N8_l2", 8, 256, 256, 32, 32, 3, film=True, relu=False, film_self=True),
    lambda: conv_case("tc_3x3_96to32_N8_l2", 8, 256, 256, 64, 32, 3, c1=32),
    lambda: conv_case("tc_deconv_64_N8_l2", 8, 128, 128, 64, 64, 1, deconv=True),
    lambda: conv_case("tc_3x3_32to32_film_N64", 64, 256, 256, 32, 32, 3, film=True, relu=False),
    lambda: conv_case("tc_3x3_32to32_filmA_N64", 64, 256, 256, 32, 32, 3, film=True, relu=False, film_self=True),
    lambda: conv_case("tc_3x3_64to64_filmA_N64", 64, 128, 128, 64, 64, 3, film=True, relu=False, film_self=True),
    lambda: conv_case("tc_3x3_96to96_filmA_N64", 64, 64, 64, 96, 96, 3, film=True, relu=False, film_self=True),
    lambda: conv_case("tc_3x3_32to32_film_pre_N32", 32, 256, 256, 32, 32, 3, film=True, relu=False, pre=True),
    lambda: conv_case("tc_3x3_32to32_mask_N32", 32, 256, 256, 32, 32, 3, mask=True, relu=False),
    lambda: conv_case("tc_3x3_32to32_head_N64", 64, 256, 256, 32, 32, 3, head=4),
    lambda: conv_case("tc_3x3_96to32_N64", 64, 256, 256, 64, 32, 3, c1=32),
    lambda: conv_case("tc_3x3_64to64_plain_N64", 64, 128, 128, 64, 64, 3),
    lambda: conv_case("tc_3x3_64to64_film_N64", 64, 128, 128, 64, 64, 3, film=True, relu=False),
    lambda: conv_case("tc_3x3_160to64_N64", 64, 128, 128, 64, 64, 3, c1=96),
    lambda: conv_case("tc_3x3_96to96_film_N64", 64, 64, 64, 96, 96, 3, film=True, relu=False),
    lambda: conv_case("tc_3x3_128to128_film_N64", 64, 32, 32, 128, 128, 3, film=True, relu=False),
    lambda: conv_case("tc_deconv_64_N64", 64, 128, 128, 64, 64, 1, deconv=True),
    lambda: conv_case("tc_deconv_96_N64", 64, 64, 64, 96, 96, 1, deconv=True),
    lambda: conv_case("tc_5x5_16to16_N96", 96, 256, 256, 16, 16, 5),
    lambda: conv_case("tc_5x5_16to16_mask_N32", 32, 256, 256, 16, 16, 5, mask=True, relu=False),
    lambda: conv_case("tc_5x5_32to32_N96", 96, 128, 128, 32, 32, 5),
    lambda: conv_case("tc_5x5_16to32_N96", 96, 128, 128, 16, 32, 5),
    lambda: conv_case("tc_5x5_32to16_mask_N96", 96, 128, 128, 32, 16, 5, mask=True, relu=False),
    lambda: conv_case("tc_5x5_32to32_mask_N96", 96, 128, 128, 32, 32, 5, mask=True, relu=False),
    lambda: conv_case("tc_5x5_16to16_mask_N96", 96, 256, 256, 16, 16, 5, mask=True, relu=False),
    lambda: conv_case("tc_3x3_32to64_N64", 64, 128, 128, 32, 64, 3),
    lambda: conv_case("tc_3x3_256to256_N96", 96, 16, 16, 256, 256, 3),
    lambda: wgrad_case("wgrad_tc_5x5_16to16_N64", 64, 256, 256, 16, 16, 5, 1),
    lambda: wgrad_case("wgrad_tc_3x3_32to32_N32", 32, 256, 256, 32, 32, 3, 1),
    lambda: wgrad_case("wgrad_tc_3x3_64to64_N32", 32, 128, 128, 64, 64, 3, 1),
    lambda: wgrad_case("wgrad_tc_5x5_32to32_N64", 64, 128, 128, 32, 32, 5, 1),
]


def main():
    sel = sys.argv[1:]
    import os
    reps = int(os.environ.get("KBENCH_REPS", "10"))
    warm = int(os.environ.get("KBENCH_WARMUP", "2"))
    print("%-34s %9s %9s %9s" % ("case", "ms", "TFLOP/s", "GB/s"))
    for mk in CASES:
        name, run, flops, nbytes, keep = mk()
        if sel and not any(s in name for s in sel):
            del keep
            continue
        try:
            for _ in range(warm):
                run()
            torch.cuda.synchronize()
            # spin the clocks up: an idle B200 sits at 120 MHz and needs tens of milliseconds of load to reach its
            # boost clock -- ten 0.1 ms launches from idle measure the ramp, not the kernel (KBENCH_SPIN_MS=0: off)
            spin_ms = float(os.environ.get("KBENCH_SPIN_MS", "0" if reps <= 1 else "60"))
            if spin_ms > 0:
                import time
                t_end = time.perf_counter() + spin_ms * 1e-3
                while time.perf_counter() < t_end:
                    for _ in range(20):
                        run()
                    torch.cuda.synchronize()
        except RuntimeError as e:
            print("%-34s unsupported by this build (%s)" % (name, str(e)[:60]))
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nrun = reps if reps <= 1 else max(reps, 50)
        for _ in range(nrun):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / nrun
        print("%-34s %9.4f %9.1f %9.1f" % (name, ms, flops / ms / 1e9, nbytes / ms / 1e6), flush=True)
        del keep, run
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
