"""Parity checks of each CUDA kernel (through the C ABI) against an fp32 torch reference fed the SAME
16-bit-rounded inputs and weights (then only accumulation order and one output rounding
differ, so fp16 can meet rtol 1e-3 per op; bf16's output rounding alone is 2^-8, so it gets rtol 1.6e-2).

Every check returns a dict(name, err, tol, ok, ...); ``tests/test_ops_gpu.py`` asserts on them and
``scripts/gpu_report.py`` prints them all without stopping at the first failure.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from rcdms_b200 import ops

# tolerances: |ours - ref| <= atol + rtol * |ref|, with atol scaled by the reference's RMS
TOL = {torch.float16: (1e-3, 1e-3), torch.bfloat16: (1.6e-2, 8e-3)}


def _result(name, out, ref, dtype, extra=None, rtol_mul=1.0):
    out = out.float()
    ref = ref.float()
    rtol, atol_rel = TOL[dtype]
    rtol *= rtol_mul
    scale = ref.pow(2).mean().sqrt().item() + 1e-12
    atol = atol_rel * rtol_mul * scale
    diff = (out - ref).abs()
    bad = diff > (atol + rtol * ref.abs())
    finite = bool(torch.isfinite(out).all())
    r = dict(name=name, dtype=str(dtype).replace("torch.", ""), max_abs=diff.max().item(), ref_rms=scale,
             rel=diff.max().item() / scale, frac_bad=bad.float().mean().item(), ok=finite and not bool(bad.any()))
    if extra:
        r.update(extra)
    return r


def _rand(shape, dtype, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(dtype)


def _gen(seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return g


def check_linear(M, N, K, dtype, bias=True, residual=False, tile_n=0, simple=False, seed=0):
    g = _gen(seed)
    a = _rand((M, K), dtype, g)
    w = _rand((N, K), dtype, g, 1.0 / math.sqrt(K))
    b = torch.randn((N,), generator=g, device="cuda") if bias else None
    r = _rand((M, N), dtype, g) if residual else None
    out = ops.linear(a, w, b, r, tile_n=tile_n, simple=simple)
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b
    if residual:
        ref = ref + r.float()
    # with a residual the result is rounded twice (GEMM output, then the add) exactly like the reference's
    # `attn(...) + hidden_states` on 16-bit tensors -> twice the single-rounding tolerance
    return _result(f"linear M{M} N{N} K{K} bn{tile_n} b{int(bias)} r{int(residual)} s{int(simple)}", out, ref, dtype,
                   rtol_mul=2.0 if residual else 1.0)


def check_geglu(M, C, dtype, simple=False, seed=1):
    g = _gen(seed)
    a = _rand((M, C), dtype, g)
    w = _rand((8 * C, C), dtype, g, 1.0 / math.sqrt(C))
    b = torch.randn((8 * C,), generator=g, device="cuda") * 0.5
    out = ops.geglu_linear(a, w, b, simple=simple)
    hg = a.float() @ w.float().t() + b
    h, gate = hg.chunk(2, dim=-1)
    ref = h * F.gelu(gate)
    return _result(f"geglu M{M} C{C} s{int(simple)}", out, ref, dtype, rtol_mul=2.0)


def check_conv3x3(n, h, w, cin, cout, stride, dtype, residual=False, simple=False, seed=2):
    g = _gen(seed)
    x = _rand((n, h, w, cin), dtype, g)
    wt = _rand((cout, cin, 3, 3), dtype, g, 1.0 / math.sqrt(9 * cin))
    b = torch.randn((cout,), generator=g, device="cuda")
    r = _rand((n, h // stride, w // stride, cout), dtype, g) if residual else None
    out = ops.conv3x3(x, wt, b, r, stride=stride, simple=simple)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=stride, padding=1).permute(0, 2, 3, 1)
    if residual:
        ref = ref + r.float()
    return _result(f"conv3x3 n{n} {h}x{w} {cin}->{cout} s{stride} r{int(residual)} simple{int(simple)}", out, ref,
                   dtype, rtol_mul=2.0 if residual else 1.0)


def check_upsample_conv3x3(n, h, w, cin, cout, dtype, seed=13):
    """Upsample3D (nearest 2x, resnet.py:65) + conv3x3 folded into four 2x2 convs vs F.interpolate + F.conv2d in fp32.
    The folded weights are sums of up to four 16-bit weights rounded once, so the result differs from the reference by
    one extra weight rounding: same x2 tolerance as the other twice-rounded ops."""
    g = _gen(seed)
    x = _rand((n, h, w, cin), dtype, g)
    wt = _rand((cout, cin, 3, 3), dtype, g, 1.0 / math.sqrt(9 * cin))
    b = torch.randn((cout,), generator=g, device="cuda")
    out = ops.upsample_conv3x3(x, wt, b)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, wt.float(), b, padding=1).permute(0, 2, 3, 1)
    return _result(f"upsample_conv3x3 n{n} {h}x{w} {cin}->{cout}", out, ref, dtype, rtol_mul=2.0)


def check_groupnorm(nstat, rows_per_stat, C, dtype, silu=True, eps=1e-5, seed=3):
    g = _gen(seed)
    x = _rand((nstat * rows_per_stat, C), dtype, g) * 2 + 0.5
    gamma = 1 + 0.2 * torch.randn((C,), generator=g, device="cuda")
    beta = 0.2 * torch.randn((C,), generator=g, device="cuda")
    out = ops.group_norm(x, gamma, beta, 32, rows_per_stat, eps, silu)
    xr = x.float().reshape(nstat, rows_per_stat, C).permute(0, 2, 1)  # (N, C, L)
    ref = F.group_norm(xr, 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(nstat * rows_per_stat, C)
    return _result(f"groupnorm nstat{nstat} rows{rows_per_stat} C{C} silu{int(silu)}", out, ref, dtype, rtol_mul=2.0)


def check_groupnorm_from_stats(n_img, hw, frames_per_stat, C0, C1, K, dtype, silu=True, eps=1e-5, residual=False, seed=11):
    """Two GEMMs write x0 [rows, C0] / x1 [rows, C1] and emit their GroupNorm chunk statistics from the epilogue; the
    streaming apply kernel normalises the virtual concat.  Reference: F.group_norm of the GEMM OUTPUTS (as stored, i.e.
    rounded) over (frames_per_stat * hw) rows per statistic batch."""
    g = _gen(seed)
    rows = n_img * hw
    outs, accs = [], []
    for C in (C0, C1):
        if C == 0:
            continue
        a = _rand((rows, K), dtype, g)
        w = _rand((C, K), dtype, g, 1.0 / math.sqrt(K))
        b = torch.randn((C,), generator=g, device="cuda") * 0.5 + 0.3
        r = _rand((rows, C), dtype, g) if residual else None
        o, acc = ops.linear_gnstats(a, w, b, r, hw)
        ref_o = a.float() @ w.float().t() + b + (r.float() if residual else 0)
        assert (o.float() - ref_o).abs().max().item() < 0.1 * ref_o.abs().max().item()
        outs.append(o)
        accs.append(acc)
    C = C0 + C1
    gamma = 1 + 0.2 * torch.randn((C,), generator=g, device="cuda")
    beta = 0.2 * torch.randn((C,), generator=g, device="cuda")
    rps = frames_per_stat * hw
    out = ops.group_norm_from_stats(outs[0], accs[0], gamma, beta, 32, rps, hw, eps, silu,
                                    outs[1] if C1 else None, accs[1] if C1 else None)
    x = torch.cat([o.float() for o in outs], dim=1)
    xr = x.reshape(rows // rps, rps, C).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(rows, C)
    return _result(f"groupnorm_from_stats img{n_img} hw{hw} fps{frames_per_stat} C{C0}+{C1} silu{int(silu)} r{int(residual)}",
                   out, ref, dtype, rtol_mul=2.0)


def check_layernorm(rows, C, dtype, pe=False, seed=4):
    g = _gen(seed)
    x = _rand((rows, C), dtype, g) * 1.5 + 0.3
    gamma = 1 + 0.2 * torch.randn((C,), generator=g, device="cuda")
    beta = 0.2 * torch.randn((C,), generator=g, device="cuda")
    frames, rpf = 5, max(1, rows // 10)
    pet = torch.randn((frames, C), generator=g, device="cuda") if pe else None
    out = ops.layer_norm(x, gamma, beta, 1e-5, pet, rpf, frames)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    if pe:
        fr = (torch.arange(rows, device="cuda") // rpf) % frames
        ref = ref + pet[fr]
    return _result(f"layernorm rows{rows} C{C} pe{int(pe)}", out, ref, dtype, rtol_mul=2.0)


def check_linear_ln(M, N, K, dtype, pe=False, geglu=False, seed=8):
    """LayerNorm folded around the GEMM vs LayerNorm (fp32) -> Linear (fp32) on the same rounded inputs."""
    g = _gen(seed)
    x = _rand((M, K), dtype, g) * 1.5 + 0.3
    w = _rand((N, K), dtype, g, 1.0 / math.sqrt(K))
    gamma = 1 + 0.2 * torch.randn((K,), generator=g, device="cuda")
    beta = 0.2 * torch.randn((K,), generator=g, device="cuda")
    bias = torch.randn((N,), generator=g, device="cuda") * 0.5 if (geglu or seed % 2 == 0) else None
    frames, rpf = 5, max(1, M // 10)
    pet = torch.randn((frames, K), generator=g, device="cuda") if pe else None
    out = ops.linear_ln(x, w, gamma, beta, bias, pet, rpf, geglu)
    ln = F.layer_norm(x.float(), (K,), gamma, beta, 1e-5)
    if pe:
        fr = (torch.arange(M, device="cuda") // rpf) % frames
        ln = ln + pet[fr]
    ref = ln @ w.float().t()
    if bias is not None:
        ref = ref + bias
    if geglu:
        h, gate = ref.chunk(2, dim=-1)
        ref = h * F.gelu(gate)
    # the folded weights (W * gamma, centred) are rounded to 16 bits once more, like the reference's 16-bit LayerNorm
    # output; the GEGLU product h * gelu(g) carries both factors' errors and the bound is on the MAX of ~1e6 elements
    return _result(f"linear_ln M{M} N{N} K{K} pe{int(pe)} geglu{int(geglu)}", out, ref, dtype,
                   rtol_mul=6.0 if geglu else 3.0)


def check_ffn_fused(M, dtype, seed=12):
    """Fused GEGLU feed-forward (one kernel) vs fp32 LayerNorm -> Linear -> GEGLU -> Linear -> + residual on the same rounded
    inputs, and vs the composition of the two separate kernels (folded-LayerNorm GEGLU GEMM, then GEMM + residual)."""
    g = _gen(seed)
    C, J = 320, 1280
    y = _rand((M, C), dtype, g) * 1.5 + 0.3
    w1 = _rand((2 * J, C), dtype, g, 1.0 / math.sqrt(C))
    b1 = torch.randn((2 * J,), generator=g, device="cuda") * 0.5
    gamma = 1 + 0.2 * torch.randn((C,), generator=g, device="cuda")
    beta = 0.2 * torch.randn((C,), generator=g, device="cuda")
    w2 = _rand((C, J), dtype, g, 1.0 / math.sqrt(J))
    b2 = torch.randn((C,), generator=g, device="cuda") * 0.5
    out = ops.ffn_geglu_ln(y, w1, b1, gamma, beta, w2, b2)
    ln = F.layer_norm(y.float(), (C,), gamma, beta, 1e-5)
    hg = ln @ w1.float().t() + b1
    hmid = (hg[:, :J] * F.gelu(hg[:, J:])).to(dtype).float()  # the kernels round the intermediate to 16 bits
    ref = hmid @ w2.float().t() + b2 + y.float()
    res = _result(f"ffn_fused M{M}", out, ref, dtype, rtol_mul=3.0)
    mid = ops.linear_ln(y, w1, gamma, beta, b1, None, 1, True)
    two = ops.linear(mid, w2, b2, y)
    d = (out.float() - two.float()).abs().max().item()
    res["max_abs_vs_two_kernels"] = d
    res["ok"] = res["ok"] and d <= 2e-2 * max(1.0, two.float().abs().max().item())
    return res


def check_rowstats(M, N, K, dtype, residual=True, seed=9):
    g = _gen(seed)
    a = _rand((M, K), dtype, g)
    w = _rand((N, K), dtype, g, 1.0 / math.sqrt(K))
    b = torch.randn((N,), generator=g, device="cuda")
    r = _rand((M, N), dtype, g) if residual else None
    out, stats = ops.linear_rowstats(a, w, b, r)
    plain = ops.linear(a, w, b, r)
    tot = stats.sum(dim=0)
    o32 = out.float()
    ref = torch.stack([o32.sum(dim=1), (o32 * o32).sum(dim=1)], dim=1)
    res = _result(f"rowstats M{M} N{N} K{K} r{int(residual)}", tot, ref, torch.float16, rtol_mul=0.1)
    res["same_output_as_plain_gemm"] = bool(torch.equal(out, plain))
    res["ok"] = res["ok"] and res["same_output_as_plain_gemm"]
    return res


def check_flash(batch, heads, sq, skv, d, dtype, simple=False, seed=5, qscale=1.0, short_kv=None):
    """short_kv = 0 | 1: library option "attn_short_kv" for this call (1 = default: key ranges <= 112 on the register-resident
    mma.sync kernel; 0: the tcgen05 flash kernel like every longer range)."""
    g = _gen(seed)
    q = _rand((batch, sq, heads * d), dtype, g, qscale)
    k = _rand((batch, skv, heads * d), dtype, g)
    v = _rand((batch, skv, heads * d), dtype, g)
    if short_kv is not None:
        from rcdms_b200 import _lib
        prev = _lib.lib().rcdm_debug_set_option(b"attn_short_kv", int(short_kv))
        try:
            out = ops.flash_attention(q, k, v, heads, simple=simple)
        finally:
            _lib.lib().rcdm_debug_set_option(b"attn_short_kv", prev)
    else:
        out = ops.flash_attention(q, k, v, heads, simple=simple)

    def split(t, s):
        return t.float().reshape(batch, s, heads, d).permute(0, 2, 1, 3)
    ref = F.scaled_dot_product_attention(split(q, sq), split(k, skv), split(v, skv))
    ref = ref.permute(0, 2, 1, 3).reshape(batch, sq, heads * d)
    return _result(f"flash b{batch} h{heads} Sq{sq} Skv{skv} d{d} simple{int(simple)} qs{qscale} short{short_kv}", out, ref,
                   dtype, rtol_mul=3.0)


def check_temporal(batch, frames, hw, heads, d, dtype, seed=6):
    g = _gen(seed)
    C = heads * d
    qkv = _rand((batch * frames * hw, 3 * C), dtype, g)
    out = ops.temporal_attention(qkv, batch, frames, hw, heads)
    t = qkv.float().reshape(batch, frames, hw, 3, heads, d)
    q, k, v = (t[:, :, :, i].permute(0, 2, 3, 1, 4) for i in range(3))  # (b, hw, heads, f, d)
    ref = F.scaled_dot_product_attention(q, k, v)  # attention over f
    ref = ref.permute(0, 3, 1, 2, 4).reshape(batch * frames * hw, C)
    return _result(f"temporal b{batch} f{frames} hw{hw} h{heads} d{d}", out, ref, dtype, rtol_mul=3.0)


def check_context_fusion(B, L, text_dim, Sv, vis_dim, hidden, heads, dtype, seed=12):
    """local_feature (fine_stack / semantic_stack, stage2_batchtest_rcdms_model.py:117-149) on the B200 kernels vs the
    reference's own torch ops (nn.Linear + nn.MultiheadAttention) in fp32 on the same rounded weights / inputs."""
    from rcdms_b200.pipelines.RCDMs_pipeline import local_feature
    torch.manual_seed(seed)
    m = local_feature(text_dim, vis_dim, hidden, heads)
    with torch.no_grad():
        for prm in m.parameters():
            if prm.dim() == 1:
                prm.normal_(0, 0.1)  # MHA biases are zero-initialised: make them count
    g = _gen(seed)
    text = _rand((B, L, text_dim), dtype, g)
    vis = _rand((B, Sv, vis_dim), dtype, g)
    mh = m.to("cuda", dtype)
    out = mh(vis, text)
    ref_m = local_feature(text_dim, vis_dim, hidden, heads).to("cuda")
    ref_m.load_state_dict({k: v.float() for k, v in mh.state_dict().items()})
    with torch.no_grad():
        ref = ref_m.forward_torch(vis.float(), text.float())
    return _result(f"context_fusion B{B} L{L} Sv{Sv} vis{vis_dim} hid{hidden}", out, ref, dtype, rtol_mul=4.0)


def check_ddim(clips, f, h, w, dtype, cfg=True, seed=7):
    from oracle.loop_ref import make_scheduler
    g = _gen(seed)
    nb = 2 * clips if cfg else clips
    eps = _rand((nb, 4, f, h, w), dtype, g)
    lat = _rand((clips, 4, f, h, w), dtype, g)
    mask = (torch.rand((clips, 1, f, h, w), generator=g, device="cuda") > 0.5).to(dtype)
    ml = _rand((clips, 4, f, h, w), dtype, g, 0.18)
    sched = make_scheduler()
    sched.set_timesteps(50)
    t = 501
    a_t = float(sched.alphas_cumprod[t])
    a_prev = float(sched.alphas_cumprod[t - 20])
    lat32 = lat.float().clone()
    out, nxt = ops.ddim_cfg_step(eps, lat32, a_t, a_prev, 2.0, cfg, dtype, mask, ml, dtype)
    # (a) bit-exact against the reference's own arithmetic: torch ops on `dtype` tensors with fp32 0-dim scalars
    #     (RCDMs_pipeline.py:493-497 + diffusers DDIMScheduler.step as restated in rcdms_b200.schedulers)
    e16 = eps
    if cfg:
        eu, ec = eps.chunk(2)
        e16 = eu + 2.0 * (ec - eu)
    ref16 = sched.step(e16, t, lat, eta=0.0).prev_sample
    exact = bool(torch.equal(out, ref16))
    # (b) close to the fp32 oracle (per-op 16-bit rounding of the reference included => looser tolerance)
    e = eps.float()
    if cfg:
        eu, ec = e.chunk(2)
        e = eu + 2.0 * (ec - eu)
    ref = sched.step(e.cpu(), t, lat.float().cpu(), eta=0.0).prev_sample.cuda()
    r = _result(f"ddim clips{clips} {f}x{h}x{w} cfg{int(cfg)}", out, ref, dtype, rtol_mul=8.0)
    r["bit_exact_vs_reference_arithmetic"] = exact
    r["ok"] = r["ok"] and exact
    lat2 = torch.cat([out] * 2) if cfg else out
    ref_next = torch.cat([lat2, torch.cat([mask] * (2 if cfg else 1)), torch.cat([ml] * (2 if cfg else 1))], dim=1)
    r["next_exact"] = bool(torch.equal(nxt, ref_next))
    r["master_matches"] = bool(torch.equal(lat32.to(dtype), out))
    r["ok"] = r["ok"] and r["next_exact"] and r["master_matches"]
    return r


def all_op_checks(dtypes=(torch.float16, torch.bfloat16), quick=False):
    """Yield thunks so a failing (or crashing) check does not hide the others."""
    for dt in dtypes:
        for (M, N, K) in [(256, 160, 64), (640, 320, 320), (1000, 128, 96), (130, 64, 128), (4096, 960, 320),
                          (2560, 1280, 1280), (10, 256, 256), (850, 640, 768)]:
            yield lambda M=M, N=N, K=K, dt=dt: check_linear(M, N, K, dt, bias=True, residual=(N % 2 == 0 and M > 200))
        for bn in (64, 128, 160, 192):
            yield lambda bn=bn, dt=dt: check_linear(512, 320, 256, dt, tile_n=bn)
        # 192-wide tiles (chosen by gemm_plain_bn where they save a wave; here forced): ragged last tile, residual
        yield lambda dt=dt: check_linear(1000, 200, 64, dt, residual=True, tile_n=192)
        yield lambda dt=dt: check_linear(2560, 3840, 1280, dt, bias=False)
        yield lambda dt=dt: check_linear(384, 192, 320, dt, bias=False)
        yield lambda dt=dt: check_linear(40960, 320, 1280, dt, residual=True)
        # shapes whose tile count does not fill the SMs evenly -> stream-K decomposition (gemm_host.cu)
        yield lambda dt=dt: check_linear(640, 1280, 5120, dt, residual=True)
        yield lambda dt=dt: check_linear(2560, 1280, 5120, dt, residual=True)
        yield lambda dt=dt: check_linear(10240, 640, 2560, dt, bias=False)
        yield lambda dt=dt: check_conv3x3(10, 16, 16, 1280, 1280, 1, dt, residual=True)
        yield lambda dt=dt: check_conv3x3(10, 8, 8, 1280, 640, 1, dt)
        yield lambda dt=dt: check_conv3x3(10, 16, 16, 640, 1280, 2, dt)
        yield lambda dt=dt: check_geglu(640, 640, dt)
        yield lambda dt=dt: check_linear(300, 320, 64, dt, simple=True)
        for (M, N, K, pe, gg) in [(640, 960, 320, False, False), (640, 960, 320, True, False), (1000, 320, 320, False, False),
                                  (2560, 3840, 1280, True, False), (640, 2560, 320, False, True),
                                  (300, 512, 64, True, True), (10240, 1920, 640, False, False)]:
            yield lambda a=(M, N, K, pe, gg), dt=dt: check_linear_ln(a[0], a[1], a[2], dt, pe=a[3], geglu=a[4])
        for (M, N, K, rs) in [(640, 320, 320, True), (1000, 640, 128, False), (2560, 1280, 1280, True), (130, 64, 64, True),
                              (300, 192, 96, True)]:
            yield lambda a=(M, N, K, rs), dt=dt: check_rowstats(a[0], a[1], a[2], dt, residual=a[3])
        for (M, C) in [(256, 64), (1024, 320), (100, 128)]:
            yield lambda M=M, C=C, dt=dt: check_geglu(M, C, dt)
        yield lambda dt=dt: check_geglu(128, 64, dt, simple=True)
        for M in (128, 1000, 20480):
            yield lambda M=M, dt=dt: check_ffn_fused(M, dt)
        for (n, h, w, cin, cout, s) in [(10, 8, 8, 64, 64, 1), (2, 64, 64, 128, 160, 1), (10, 4, 4, 128, 64, 1),
                                        (10, 2, 2, 64, 128, 1), (10, 1, 1, 256, 256, 1), (3, 16, 16, 320, 640, 1),
                                        (10, 32, 32, 64, 4, 1), (10, 8, 8, 64, 128, 2), (2, 64, 64, 64, 64, 2),
                                        (10, 2, 2, 128, 128, 2), (5, 16, 16, 192, 320, 2)]:
            yield lambda a=(n, h, w, cin, cout, s), dt=dt: check_conv3x3(*a, dt, residual=(a[4] % 8 == 0))
        for (n, h, w, cin, cout) in [(10, 8, 8, 1280, 1280), (10, 16, 16, 1280, 1280), (2, 32, 32, 640, 640), (3, 4, 8, 64, 128),
                                     (10, 2, 2, 64, 64)]:
            yield lambda a=(n, h, w, cin, cout), dt=dt: check_upsample_conv3x3(*a, dt)
        yield lambda dt=dt: check_conv3x3(4, 8, 8, 64, 64, 1, dt, simple=True)
        yield lambda dt=dt: check_conv3x3(4, 8, 8, 64, 64, 2, dt, simple=True)
        for (ns, rps, C, silu) in [(2, 5 * 64, 320, True), (10, 64, 64, False), (2, 5 * 4096, 320, True),
                                   (10, 1, 256, True), (2, 5 * 256, 960, True), (10, 1024, 640, False),
                                   (2, 20, 2560, True), (2, 5 * 4096, 960, True), (3, 1000, 192, False),
                                   (16, 4096, 320, False), (160, 64, 1280, True), (1, 7, 64, True)]:
            yield lambda a=(ns, rps, C), silu=silu, dt=dt: check_groupnorm(*a, dt, silu=silu,
                                                                           eps=1e-5 if silu else 1e-6)
        # statistics from the producing GEMMs' epilogues + streaming apply: (images, hw, frames per statistic, C0, C1, K)
        #   resnet norms span the 5 frames (and an un-materialised skip concat whose groups straddle the two tensors:
        #   1280 + 640 -> groups of 60, 640 + 320 -> groups of 30); transformer / motion norms are per frame
        for (ni, hw, fps, c0, c1, K, silu, rs) in [(10, 64, 5, 320, 0, 64, True, False), (10, 256, 1, 320, 0, 320, False, True),
                                                  (10, 1024, 5, 640, 320, 64, True, True), (10, 64, 5, 1280, 640, 128, True, False),
                                                  (10, 4096, 5, 320, 320, 64, True, False), (10, 4096, 1, 320, 0, 64, False, True),
                                                  (20, 64, 5, 1280, 1280, 64, True, True), (10, 256, 1, 1280, 0, 64, False, False)]:
            yield lambda a=(ni, hw, fps, c0, c1, K), silu=silu, rs=rs, dt=dt: check_groupnorm_from_stats(
                *a, dt, silu=silu, eps=1e-5 if silu else 1e-6, residual=rs)
        for (rows, C, pe) in [(640, 320, False), (640, 320, True), (100, 64, True), (2560, 1280, True),
                              (1000, 640, False), (77, 256, False)]:
            yield lambda a=(rows, C, pe), dt=dt: check_layernorm(a[0], a[1], dt, pe=a[2])
        for (b, hds, sq, skv, d) in [(4, 8, 256, 256, 40), (3, 8, 64, 85, 40), (2, 8, 1024, 1024, 80),
                                     (2, 8, 256, 256, 160), (10, 8, 64, 64, 8), (10, 8, 16, 7, 16),
                                     (10, 8, 4, 4, 32), (10, 8, 1, 1, 32), (2, 2, 4096, 4096, 40),
                                     (2, 8, 1024, 91, 80), (2, 4, 300, 200, 40),
                                     # odd numbers of 64-key tiles (the MMA loop is unrolled by stage parity), ragged last tiles,
                                     # a head dim whose padded chunk is not the 6th (d = 24 -> 32)
                                     (2, 8, 256, 150, 40), (1, 4, 128, 320, 80), (2, 8, 130, 450, 160), (2, 2, 200, 129, 24)]:
            yield lambda a=(b, hds, sq, skv, d), dt=dt: check_flash(*a, dt)
        # short key ranges (cross-attention to 85 / 91 context tokens at the UNet's four levels, the 8x8 level's self-attention,
        # the 112-key limit, ragged query tiles): the register-resident mma.sync kernel (default) AND the flash kernel on the same problems
        for (b, hds, sq, skv, d) in [(10, 8, 4096, 85, 40), (10, 8, 1024, 85, 80), (10, 8, 256, 85, 160), (10, 8, 64, 85, 160),
                                     (16, 8, 1024, 91, 80), (10, 8, 64, 64, 160), (2, 8, 100, 112, 40), (3, 4, 33, 5, 64),
                                     (2, 8, 256, 113, 40)]:
            for sk in (1, 0):
                yield lambda a=(b, hds, sq, skv, d), dt=dt, sk=sk: check_flash(*a, dt, short_kv=sk)
        yield lambda dt=dt: check_flash(2, 4, 512, 512, 40, dt, qscale=8.0)
        yield lambda dt=dt: check_flash(2, 8, 64, 85, 40, dt, qscale=8.0, short_kv=1)
        yield lambda dt=dt: check_flash(2, 8, 64, 85, 40, dt, simple=True)
        for (b, f, hw, hds, d) in [(2, 5, 64, 8, 40), (2, 5, 16, 8, 8), (2, 5, 4096, 8, 40), (2, 5, 1, 8, 32),
                                   (2, 5, 256, 8, 160), (1, 3, 10, 8, 80)]:
            yield lambda a=(b, f, hw, hds, d), dt=dt: check_temporal(*a, dt)
        # context fusion: PororoSV / FlintstonesSV shapes (257 CLIP patch tokens of width 1664; one 1280-wide embedding)
        yield lambda dt=dt: check_context_fusion(2, 85, 768, 257, 1664, 768, 8, dt)
        yield lambda dt=dt: check_context_fusion(8, 91, 768, 1, 1280, 768, 8, dt)
        yield lambda dt=dt: check_context_fusion(3, 7, 96, 9, 16, 96, 8, dt)   # d = 12: the library's CUDA-core attention
        yield lambda dt=dt: check_context_fusion(4, 7, 96, 1, 12, 96, 8, dt)   # K = 12 zero-padded to 16
        yield lambda dt=dt: check_ddim(1, 5, 64, 64, dt, cfg=True)
        yield lambda dt=dt: check_ddim(3, 5, 8, 8, dt, cfg=False)
