"""GPU bring-up script (not a pytest): exercises every C-ABI entry point against the oracle, prints
per-case errors and first timings, and writes gpurun_out/first_light.json.  Sections are
independent so one failing kernel does not hide the others.

    gpurun -- 'timeout 600 python tests/gpu_first_light.py [section ...]'
"""
from __future__ import annotations

import json
import os
import sys
import time
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from mlx_cuda_distributed_pretraining_b200 import ops  # noqa: E402
from mlx_cuda_distributed_pretraining_b200._lib import lib  # noqa: E402
from oracle import reference_math as R  # noqa: E402

OUT = {}
DEV = "cuda"


def rel_fro(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def time_cuda(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


def section(name):
    def deco(fn):
        fn._section = name
        return fn
    return deco


@section("gemm")
def sec_gemm():
    res = []
    torch.manual_seed(0)
    cases = [
        # (a_mn, b_mn, M, N, K, batch, bn)
        (0, 0, 128, 256, 64, 1, 0), (0, 0, 128, 256, 256, 1, 0), (0, 0, 256, 512, 512, 2, 256),
        (0, 0, 256, 512, 512, 2, 128), (0, 0, 200, 328, 136, 3, 0), (0, 0, 1024, 1024, 1024, 1, 0),
        (0, 1, 128, 256, 64, 1, 0), (0, 1, 256, 512, 512, 2, 256), (0, 1, 256, 512, 512, 2, 128),
        (0, 1, 200, 328, 136, 3, 0),
        (1, 1, 128, 256, 64, 1, 0), (1, 1, 256, 512, 512, 2, 256), (1, 1, 256, 512, 1000, 2, 128),
        (1, 1, 200, 328, 131, 3, 0),
        (1, 0, 256, 512, 512, 2, 0),
    ]
    for (a_mn, b_mn, M, N, K, batch, bn) in cases:
        a = torch.randn((batch, K, M) if a_mn else (batch, M, K), device=DEV).to(torch.bfloat16)
        b = torch.randn((batch, K, N) if b_mn else (batch, N, K), device=DEV).to(torch.bfloat16)
        af = a.float().transpose(1, 2) if a_mn else a.float()
        bf = b.float() if b_mn else b.float().transpose(1, 2)
        ref = torch.bmm(af, bf)
        try:
            out = ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.float32, force_bn=bn)
            torch.cuda.synchronize()
            err = rel_fro(out, ref)
        except Exception as e:  # noqa: BLE001
            err = f"EXC {e}"
        res.append({"case": [a_mn, b_mn, M, N, K, batch, bn], "rel_fro": err})
        print("gemm", a_mn, b_mn, M, N, K, batch, bn, "->", err, flush=True)
    # epilogue: alpha/beta/C, per-batch scalars, bf16 out
    M, N, K, batch = 256, 512, 320, 3
    a = torch.randn(batch, M, K, device=DEV).to(torch.bfloat16)
    b = torch.randn(batch, N, K, device=DEV).to(torch.bfloat16)
    c = torch.randn(batch, M, N, device=DEV).to(torch.bfloat16)
    av = torch.rand(batch, device=DEV) + 0.5
    bv = torch.rand(batch, device=DEV) + 0.5
    out = ops.gemm(a, b, c=c, alpha=0.5, beta=-1.25, alpha_vec=av, beta_vec=bv)
    ref = 0.5 * av[:, None, None] * torch.bmm(a.float(), b.float().transpose(1, 2)) - 1.25 * bv[:, None, None] * c.float()
    err = rel_fro(out, ref)
    print("gemm epilogue bf16 ->", err, flush=True)
    res.append({"case": "epilogue_bf16", "rel_fro": err})
    c32 = torch.randn(batch, M, N, device=DEV)
    out = ops.gemm(a, b, c=c32, alpha=2.0, beta=0.75, out_dtype=torch.float32)
    ref = 2.0 * torch.bmm(a.float(), b.float().transpose(1, 2)) + 0.75 * c32
    err = rel_fro(out, ref)
    print("gemm epilogue f32 ->", err, flush=True)
    res.append({"case": "epilogue_f32", "rel_fro": err})
    OUT["gemm"] = res


@section("gemm_perf")
def sec_gemm_perf():
    res = []
    for (M, N, K, batch, bn) in [(4096, 4096, 4096, 1, 256), (8192, 8192, 8192, 1, 256),
                                 (1024, 1024, 1024, 24, 256), (1024, 1024, 1024, 24, 128),
                                 (1024, 1024, 2816, 12, 256)]:
        a = torch.randn(batch, M, K, device=DEV).to(torch.bfloat16)
        b = torch.randn(batch, N, K, device=DEV).to(torch.bfloat16)
        out = torch.empty(batch, M, N, device=DEV, dtype=torch.bfloat16)
        ms = time_cuda(lambda: ops.gemm(a, b, out=out, force_bn=bn))
        tf = 2.0 * M * N * K * batch / ms / 1e9
        ms_t = time_cuda(lambda: torch.bmm(a, b.transpose(1, 2)))
        tf_t = 2.0 * M * N * K * batch / ms_t / 1e9
        print(f"gemm_perf M{M} N{N} K{K} b{batch} bn{bn}: {ms:.3f} ms {tf:.1f} TF | cuBLAS {ms_t:.3f} ms {tf_t:.1f} TF", flush=True)
        res.append({"shape": [M, N, K, batch, bn], "ms": ms, "tflops": tf, "cublas_tflops": tf_t})
    OUT["gemm_perf"] = res


@section("proj_perf")
def sec_proj_perf():
    """Library GEMM vs the hand-written GEMM on every projection shape of a C2 layer (T = 16384 tokens):
    forward y = x W^T, dgrad dx = dy W, wgrad dW = dy^T x."""
    T = 16384
    res = []
    for name, n_out, n_in in [("q/o", 1024, 1024), ("k/v", 512, 1024), ("gate/up", 2816, 1024), ("down", 1024, 2816)]:
        x = torch.randn(T, n_in, device=DEV).to(torch.bfloat16)
        w = (torch.randn(n_out, n_in, device=DEV) * 0.02).to(torch.bfloat16)
        dy = torch.randn(T, n_out, device=DEV).to(torch.bfloat16)
        y = torch.empty(T, n_out, device=DEV, dtype=torch.bfloat16)
        dx = torch.empty(T, n_in, device=DEV, dtype=torch.bfloat16)
        dw = torch.zeros(n_out, n_in, device=DEV, dtype=torch.bfloat16)
        fl = 2.0 * T * n_out * n_in
        row = {"proj": name}
        for tag, lib_fn, my_fn in [
            ("fwd", lambda: torch.nn.functional.linear(x, w), lambda: ops.gemm(x, w, out=y.unsqueeze(0))),
            ("dgrad", lambda: torch.matmul(dy, w), lambda: ops.gemm(dy, w, b_mn=True, out=dx.unsqueeze(0))),
            ("wgrad", lambda: dw.addmm_(dy.t(), x),
             lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, c=dw.unsqueeze(0), beta=1.0, out=dw.unsqueeze(0)))]:
            ms_l, ms_m = time_cuda(lib_fn), time_cuda(my_fn)
            row[tag] = {"lib_us": ms_l * 1e3, "mine_us": ms_m * 1e3, "lib_tf": fl / ms_l / 1e9, "mine_tf": fl / ms_m / 1e9}
        print(f"proj_perf {name:8s} " + " | ".join(
            f"{t}: lib {row[t]['lib_us']:.1f} us ({row[t]['lib_tf']:.0f} TF) mine {row[t]['mine_us']:.1f} us ({row[t]['mine_tf']:.0f} TF)"
            for t in ("fwd", "dgrad", "wgrad")), flush=True)
        res.append(row)
    OUT["proj_perf"] = res


NS_SHAPES_C2 = [(24, 1024, 1024), (24, 512, 1024), (24, 2816, 1024), (12, 1024, 2816), (1, 32003, 1024)]


def ns_flops(batch, r, c, steps=5):
    m, n = min(r, c), max(r, c)
    return batch * steps * (4.0 * m * m * n + 2.0 * m ** 3)


@section("ns")
def sec_ns():
    res = []
    torch.manual_seed(1)
    for (batch, r, c) in [(2, 256, 512), (2, 512, 256), (1, 1024, 1024), (2, 512, 1024), (1, 2816, 1024),
                          (1, 1024, 2816), (1, 1000, 256), (1, 32003, 1024)]:
        g = torch.randn(batch, r, c) * 0.02
        x = ops.zeropower_via_newtonschulz5(g.to(DEV))
        torch.cuda.synchronize()
        ref64 = R.newton_schulz5(g.double())
        ref_bf = R.newton_schulz5(g.float(), operand_dtype=torch.bfloat16)
        e64, ebf = rel_fro(x, ref64), rel_fro(x, ref_bf)
        sv = torch.linalg.svdvals(x[0].float().cpu()) if max(r, c) <= 4096 else None
        print(f"ns batch{batch} {r}x{c}: rel_fro vs fp64 {e64:.3e} | vs bf16-emulated {ebf:.3e}"
              + (f" | sv [{sv.min():.3f},{sv.max():.3f}]" if sv is not None else ""), flush=True)
        res.append({"shape": [batch, r, c], "rel_fro_fp64": e64, "rel_fro_bf16emu": ebf})
    OUT["ns"] = res


@section("ns_perf")
def sec_ns_perf():
    res = []
    total_ms, total_fl = 0.0, 0.0
    for (batch, r, c) in NS_SHAPES_C2:
        g = (torch.randn(batch, r, c, device=DEV) * 0.02).to(torch.bfloat16)
        inv, inv2 = ops.ns_scales(ops.sumsq(g))
        ws = torch.empty(ops.ns_workspace_bytes(batch, r, c, 5), device=DEV, dtype=torch.uint8)
        xo = torch.empty_like(g)
        ms = time_cuda(lambda: ops.newton_schulz_raw(g, xo, inv, inv2, ws, 5), warm=2, iters=5)
        fl = ns_flops(batch, r, c)
        print(f"ns_perf batch{batch} {r}x{c}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOPS", flush=True)
        res.append({"shape": [batch, r, c], "ms": ms, "tflops": fl / ms / 1e9})
        total_ms += ms
        total_fl += fl
    print(f"ns_perf C2 total: {total_ms:.3f} ms, {total_fl / 1e12:.3f} TF -> {total_fl / total_ms / 1e9:.1f} TFLOPS", flush=True)
    OUT["ns_perf"] = {"cases": res, "total_ms": total_ms, "tflops": total_fl / total_ms / 1e9}


@section("elementwise")
def sec_elementwise():
    res = {}
    torch.manual_seed(2)
    batch, r, c = 3, 64, 136
    g = torch.randn(batch, r, c, device=DEV).to(torch.bfloat16)
    buf = torch.randn(batch, r, c, device=DEV)
    buf0 = buf.clone()
    u = torch.empty(batch, r, c, device=DEV, dtype=torch.bfloat16)
    ss = torch.empty(batch, device=DEV)
    ops.muon_momentum(g, buf, u, ss, 0.95, True, 0.5)
    gf = g.float() * 0.5
    bref = 0.05 * gf + 0.95 * buf0
    uref = gf + 0.95 * bref
    res["momentum_buf"] = rel_fro(buf, bref)
    res["momentum_u"] = rel_fro(u, uref)
    res["momentum_sumsq"] = rel_fro(ss, (uref ** 2).sum(dim=(1, 2)))
    inv, inv2 = ops.ns_scales(ss, 1e-7)
    res["ns_scales"] = rel_fro(inv, 1 / (ss.sqrt() + 1e-7))
    # axpy
    p32 = torch.randn(1000 * 8 + 3, device=DEV)
    p0 = p32.clone()
    p16 = torch.empty_like(p32, dtype=torch.bfloat16)
    x = torch.randn_like(p32).to(torch.bfloat16)
    ops.axpy_update(p32, p16, x, -0.1)
    res["axpy_p32"] = rel_fro(p32, p0 - 0.1 * x.float())
    res["axpy_p16"] = rel_fro(p16, (p0 - 0.1 * x.float()).to(torch.bfloat16))
    # adamw
    n = 4099
    p = torch.randn(n, device=DEV); p_0 = p.clone()
    gg = torch.randn(n, device=DEV)
    m = torch.randn(n, device=DEV) * 0.1; v = torch.rand(n, device=DEV) * 0.1
    m0, v0 = m.clone(), v.clone()
    ops.adamw(p, None, gg, m, v, 1e-2, 0.9, 0.95, 1e-8, 0.1)
    mr = 0.9 * m0 + 0.1 * gg; vr = 0.95 * v0 + 0.05 * gg * gg
    pr = p_0 * (1 - 1e-2 * 0.1) - 1e-2 * mr / (vr.sqrt() + 1e-8)
    res["adamw_p"] = rel_fro(p, pr); res["adamw_m"] = rel_fro(m, mr); res["adamw_v"] = rel_fro(v, vr)
    # sgd momentum
    p = torch.randn(n, device=DEV); p_0 = p.clone(); bufn = torch.randn(n, device=DEV); b0 = bufn.clone()
    ops.sgd_momentum(p, None, gg, bufn, 0.95, True, 0.01)
    br = 0.05 * gg + 0.95 * b0
    res["sgd_p"] = rel_fro(p, p_0 - 0.01 * (gg + 0.95 * br))
    # clip_accum
    acc = torch.randn(n, device=DEV); a0 = acc.clone()
    g16 = (torch.randn(n, device=DEV) * 2).to(torch.bfloat16)
    ops.clip_accum(g16, acc, 1.0, 0.125, False)
    res["clip_accum"] = rel_fro(acc, a0 + g16.float().clamp(-1, 1) * 0.125)
    ops.clip_accum(g16, acc, 0.0, 0.5, True)
    res["clip_accum_init"] = rel_fro(acc, g16.float() * 0.5)
    # split
    src = torch.randn(300, 200, device=DEV)
    hi = torch.empty(128, 96, device=DEV, dtype=torch.bfloat16); lo = torch.empty_like(hi)
    ops.split_bf16(src, 128, 96, hi, lo, scale=0.5, diag_add=2.0)
    want = (src[:128, :96] + 2.0 * torch.eye(128, 96, device=DEV)) * 0.5
    res["split_hi_lo"] = rel_fro(hi.float() + lo.float(), want)
    for k, v_ in res.items():
        print(f"elementwise {k}: {v_:.3e}", flush=True)
    OUT["elementwise"] = res


@section("norm_rope")
def sec_norm_rope():
    res = {}
    torch.manual_seed(3)
    for dt in (torch.bfloat16, torch.float32):
        for H in (128, 1024, 2048):
            x = torch.randn(4, 37, H, device=DEV).to(dt).requires_grad_(True)
            w = (torch.rand(H, device=DEV) + 0.5).to(dt).requires_grad_(True)
            y = ops.rmsnorm(x, w, 1e-5)
            dy = torch.randn_like(y)
            y.backward(dy)
            xr = x.detach().float().cpu().requires_grad_(True)
            wr = w.detach().float().cpu().requires_grad_(True)
            yr = R.rmsnorm(xr, wr, 1e-5)
            yr.backward(dy.float().cpu())
            tag = f"{str(dt).split('.')[-1]}_H{H}"
            res[f"rms_y_{tag}"] = rel_fro(y, yr)
            res[f"rms_dx_{tag}"] = rel_fro(x.grad, xr.grad)
            res[f"rms_dw_{tag}"] = rel_fro(w.grad, wr.grad)
    for D in (16, 64, 128):
        x = torch.randn(2, 33, 4, D, device=DEV).to(torch.bfloat16).requires_grad_(True)
        cos_t, sin_t = ops.rope_tables(33, D, 10000.0, DEV)
        y = ops.rope(x, cos_t, sin_t)
        dy = torch.randn_like(y)
        y.backward(dy)
        xr = x.detach().float().cpu().requires_grad_(True)
        yr = R.rope(xr, 10000.0)
        yr.backward(dy.float().cpu())
        res[f"rope_y_D{D}"] = rel_fro(y, yr)
        res[f"rope_dx_D{D}"] = rel_fro(x.grad, xr.grad)
    for k, v_ in res.items():
        print(f"norm_rope {k}: {v_:.3e}", flush=True)
    OUT["norm_rope"] = res


def attn_case(B, S, H, Hk, D, causal, check_bwd=True):
    torch.manual_seed(4)
    q = torch.randn(B, S, H, D, device=DEV).to(torch.bfloat16).requires_grad_(True)
    k = torch.randn(B, S, Hk, D, device=DEV).to(torch.bfloat16).requires_grad_(True)
    v = torch.randn(B, S, Hk, D, device=DEV).to(torch.bfloat16).requires_grad_(True)
    scale = D ** -0.5
    o = ops.attention(q, k, v, scale, causal)
    torch.cuda.synchronize()
    qr, kr, vr = (t.detach().float().cpu().requires_grad_(True) for t in (q, k, v))
    mask = R.causal_mask(S) if causal else None
    orf = R.attention(qr, kr, vr, scale, mask)
    out = {"o": rel_fro(o, orf)}
    if check_bwd:
        do = torch.randn_like(o)
        o.backward(do)
        torch.cuda.synchronize()
        orf.backward(do.float().cpu())
        out.update(dq=rel_fro(q.grad, qr.grad), dk=rel_fro(k.grad, kr.grad), dv=rel_fro(v.grad, vr.grad))
    return out


@section("attn")
def sec_attn():
    res = []
    bwd = os.environ.get("B200_ATTN_BWD", "1") == "1"
    for (B, S, H, Hk, D, causal) in [(1, 128, 1, 1, 64, True), (1, 256, 2, 2, 64, True), (2, 512, 4, 2, 64, True),
                                     (1, 384, 4, 1, 128, True), (2, 16, 4, 2, 32, True), (1, 200, 2, 2, 64, True),
                                     (1, 256, 2, 1, 64, False), (2, 1024, 16, 8, 64, True)]:
        try:
            r = attn_case(B, S, H, Hk, D, causal, bwd)
        except Exception as e:  # noqa: BLE001
            r = {"error": str(e)}
            traceback.print_exc()
        print("attn", (B, S, H, Hk, D, causal), r, flush=True)
        res.append({"case": [B, S, H, Hk, D, causal], **r})
    OUT["attn"] = res


@section("attn_perf")
def sec_attn_perf():
    res = []
    for (B, S, H, Hk, D) in [(16, 1024, 16, 8, 64), (16, 2048, 16, 8, 64), (32, 2048, 16, 16, 128)]:
        q = torch.randn(B, S, H, D, device=DEV).to(torch.bfloat16)
        k = torch.randn(B, S, Hk, D, device=DEV).to(torch.bfloat16)
        v = torch.randn(B, S, Hk, D, device=DEV).to(torch.bfloat16)
        scale = D ** -0.5
        ms_f = time_cuda(lambda: ops.attention_fwd_raw(q, k, v, scale, True))
        o, lse = ops.attention_fwd_raw(q, k, v, scale, True)
        do = torch.randn_like(o)
        fl = 4.0 * B * H * S * S * D
        row = {"shape": [B, S, H, Hk, D], "fwd_ms": ms_f, "fwd_tflops_full": fl / ms_f / 1e9}
        if os.environ.get("B200_ATTN_BWD", "1") == "1":
            ms_b = time_cuda(lambda: ops.attention_bwd_raw(q, k, v, o, do, lse, scale, True))
            row.update(bwd_ms=ms_b, bwd_tflops_full=2.5 * fl / ms_b / 1e9)
        print("attn_perf", row, flush=True)
        res.append(row)
    OUT["attn_perf"] = res


def main():
    want = sys.argv[1:]
    print("device:", torch.cuda.get_device_name(0), "device_ok:", lib().b200_device_ok(), flush=True)
    secs = [f for f in globals().values() if callable(f) and hasattr(f, "_section")]
    for f in secs:
        if want and f._section not in want:
            continue
        t0 = time.time()
        try:
            f()
        except Exception as e:  # noqa: BLE001
            traceback.print_exc()
            OUT[f._section] = {"error": repr(e)}
        print(f"== section {f._section} done in {time.time() - t0:.1f}s", flush=True)
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    name = "first_light" + ("_" + "_".join(want) if want else "") + ".json"
    (out_dir / name).write_text(json.dumps(OUT, indent=1, default=str))


if __name__ == "__main__":
    main()
